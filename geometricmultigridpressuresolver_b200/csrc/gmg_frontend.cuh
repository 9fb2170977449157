// gmg_frontend.cuh -- the steps either side of the solve: the pointwise builders of
// HDK_GeometricFreeSurfacePressureSolver.cpp that turn the simulation's fields into the solver's inputs (domain labels,
// ghost-fluid boundary weights, cut-cell right-hand side, warm start) and its output back into them (pressure, velocity
// update).  All HBM-bound pointwise / one-neighbour kernels on the BASE grid; cell fields are x-fastest [rz][ry][rx], the face
// field of axis a has one more entry along a, SIM_RawField values are fpreal32, the arithmetic is double (GFS.h:18-19).
// Material labels: HDK_Utilities.h:17 { SOLID = 0, LIQUID = 1, AIR = 2 }; VALID_FACE = 1 (HDK_Utilities.h:21).
#pragma once

#include "gmg_common.cuh"

namespace gmg
{
constexpr int MAT_SOLID = 0, MAT_LIQUID = 1, MAT_AIR = 2;

struct BaseBox
{
    long long r[3];  // base resolution
    long long e[3];  // expanded resolution
    long long o[3];  // offset of the base grid in the expanded one
};

// HDK_Utilities.h:25-42 computeGhostFluidWeight, then the clamp of GFS.cpp:853 / :1118
__device__ __forceinline__ double ghostFluidTheta(double phi0, double phi1)
{
    double theta = 0;
    if (phi0 < 0)
    {
	if (phi1 < 0) theta = 1;
	else if (phi1 >= 0) theta = phi0 / (phi0 - phi1);
    }
    else if (phi1 < 0) theta = phi1 / (phi1 - phi0);
    return fmin(fmax(theta, .01), 1.);
}

// GFS.cpp:746-793 buildMGDomainLabels
__global__ void __launch_bounds__(BLOCK) k_fe_domain_labels(int32_t *labels, const int32_t *material, long long n)
{
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const int m = material[i];
    labels[i] = m == MAT_LIQUID ? L_INTERIOR : (m == MAT_AIR ? L_DIRICHLET : L_EXTERIOR);
}

// GFS.cpp:796-865 buildMGBoundaryWeights, one axis
__global__ void __launch_bounds__(BLOCK) k_fe_boundary_weights(double *weights, const float *cutCell, const float *liquidSurface, const float *validFaces,
							      const int32_t *labels, BaseBox g, int axis)
{
    long long fr[3] = {g.r[0], g.r[1], g.r[2]};
    ++fr[axis];
    const long long f = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (f >= fr[0] * fr[1] * fr[2]) return;
    double out = 0;
    if (validFaces[f] == 1.0f)
    {
	long long c[3] = {f % fr[0], (f / fr[0]) % fr[1], f / (fr[0] * fr[1])};
	double weight = cutCell[f];
	const bool hasB = c[axis] - 1 >= 0, hasF = c[axis] < g.r[axis];
	const long long fi = c[0] + g.r[0] * (c[1] + g.r[1] * c[2]);
	const long long stride = axis == 0 ? 1 : (axis == 1 ? g.r[0] : g.r[0] * g.r[1]);
	const long long bi = fi - stride;
	const int bl = hasB ? labels[bi] : L_EXTERIOR, fl = hasF ? labels[fi] : L_EXTERIOR;
	if ((bl == L_INTERIOR && fl == L_DIRICHLET) || (bl == L_DIRICHLET && fl == L_INTERIOR))
	    weight /= ghostFluidTheta(double(liquidSurface[bi]), double(liquidSurface[fi]));
	out = weight;
    }
    weights[f] = out;
}

// GFS.cpp:868-943 buildRHS (+ GFS.cpp:946-997 applyOldPressure when pressure != nullptr): one thread per base cell; rhs and
// solution are EXPANDED grids, written on LIQUID cells only
struct FeFields
{
    const float *velocity[3], *cutCell[3], *solidVelocity[3];
};
__global__ void __launch_bounds__(BLOCK) k_fe_rhs(double *rhs, const int32_t *material, FeFields fl, BaseBox g)
{
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= g.r[0] * g.r[1] * g.r[2]) return;
    if (material[i] != MAT_LIQUID) return;
    const long long c[3] = {i % g.r[0], (i / g.r[0]) % g.r[1], i / (g.r[0] * g.r[1])};
    double divergence = 0;
#pragma unroll
    for (int axis = 0; axis < 3; ++axis)
#pragma unroll
	for (int direction = 0; direction < 2; ++direction)
	{
	    long long fr[3] = {g.r[0], g.r[1], g.r[2]};
	    ++fr[axis];
	    long long fc[3] = {c[0], c[1], c[2]};
	    fc[axis] += direction;
	    const long long f = fc[0] + fr[0] * (fc[1] + fr[1] * fc[2]);
	    const double sign = direction == 0 ? 1. : -1.;
	    const double weight = fl.cutCell[axis][f];
	    if (weight > 0) divergence += sign * weight * double(fl.velocity[axis][f]);
	    if (fl.solidVelocity[axis] && weight < 1) divergence += sign * (1. - weight) * double(fl.solidVelocity[axis][f]);
	}
    rhs[(c[0] + g.o[0]) + g.e[0] * ((c[1] + g.o[1]) + g.e[1] * (c[2] + g.o[2]))] = divergence;
}
__global__ void __launch_bounds__(BLOCK) k_fe_old_pressure(double *solution, const float *pressure, const int32_t *material, BaseBox g)
{
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= g.r[0] * g.r[1] * g.r[2]) return;
    if (material[i] != MAT_LIQUID) return;
    const long long c[3] = {i % g.r[0], (i / g.r[0]) % g.r[1], i / (g.r[0] * g.r[1])};
    solution[(c[0] + g.o[0]) + g.e[0] * ((c[1] + g.o[1]) + g.e[1] * (c[2] + g.o[2]))] = pressure[i];
}
// GFS.cpp:1000-1047 applySolutionToPressure
__global__ void __launch_bounds__(BLOCK) k_fe_solution_to_pressure(float *pressure, const double *solution, const int32_t *material, BaseBox g)
{
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= g.r[0] * g.r[1] * g.r[2]) return;
    if (material[i] != MAT_LIQUID) return;
    const long long c[3] = {i % g.r[0], (i / g.r[0]) % g.r[1], i / (g.r[0] * g.r[1])};
    pressure[i] = float(solution[(c[0] + g.o[0]) + g.e[0] * ((c[1] + g.o[1]) + g.e[1] * (c[2] + g.o[2]))]);
}
// GFS.cpp:1050-1131 applyPressureGradient, one axis
__global__ void __launch_bounds__(BLOCK) k_fe_pressure_gradient(float *velocity, const float *liquidSurface, const float *pressure, const float *validFaces,
							       const int32_t *material, BaseBox g, int axis)
{
    long long fr[3] = {g.r[0], g.r[1], g.r[2]};
    ++fr[axis];
    const long long f = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (f >= fr[0] * fr[1] * fr[2]) return;
    if (validFaces[f] != 1.0f) return;
    const long long c[3] = {f % fr[0], (f / fr[0]) % fr[1], f / (fr[0] * fr[1])};
    if (c[axis] - 1 < 0 || c[axis] >= g.r[axis]) return;
    const long long fi = c[0] + g.r[0] * (c[1] + g.r[1] * c[2]);
    const long long stride = axis == 0 ? 1 : (axis == 1 ? g.r[0] : g.r[0] * g.r[1]);
    const long long bi = fi - stride;
    const int bm = material[bi], fm = material[fi];
    double gradient = double(pressure[fi] - pressure[bi]);  // fpreal32 - fpreal32 (GFS.cpp:1095)
    if (bm != MAT_LIQUID || fm != MAT_LIQUID) gradient /= ghostFluidTheta(double(liquidSurface[bi]), double(liquidSurface[fi]));
    velocity[f] = float(double(velocity[f]) - gradient);
}

// HDK_Utilities.cpp:87-148 buildMaterialCellLabels with isCellLiquid (:5-45) inlined: a cell with no open face (every one of its six
// cut-cell weights <= 0) is SOLID; otherwise LIQUID if its surface value is <= 0, or if the solid sample is >= 0 (:24) and an open
// face leads to an in-range neighbour whose surface value is <= 0 (:26-41); else AIR.  solidAtCentres is the solid SDF sampled at
// the surface field's cell centres (solidSurface.getValue(pos), :22-24 -- the HDK's own interpolation stays with the caller; for a
// collision field aligned with the surface field it is the field itself).
struct FeCut
{
    const float *cutCell[3];
};
__global__ void __launch_bounds__(BLOCK) k_fe_material_labels(int32_t *material, const float *liquidSurface, const float *solidAtCentres, FeCut fc, BaseBox g)
{
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= g.r[0] * g.r[1] * g.r[2]) return;
    const long long c[3] = {i % g.r[0], (i / g.r[0]) % g.r[1], i / (g.r[0] * g.r[1])};
    const long long stride[3] = {1, g.r[0], g.r[0] * g.r[1]};
    bool open[3][2];
    bool inFluid = false;
#pragma unroll
    for (int axis = 0; axis < 3; ++axis)
#pragma unroll
	for (int direction = 0; direction < 2; ++direction)
	{
	    long long fr[3] = {g.r[0], g.r[1], g.r[2]};
	    ++fr[axis];
	    long long f[3] = {c[0], c[1], c[2]};
	    f[axis] += direction;  // cellToFaceMap
	    open[axis][direction] = fc.cutCell[axis][f[0] + fr[0] * (f[1] + fr[1] * f[2])] > 0;
	    inFluid |= open[axis][direction];
	}
    int label = MAT_SOLID;
    if (inFluid)
    {
	bool liquid = liquidSurface[i] <= 0.;
	if (!liquid && solidAtCentres[i] >= 0)
	{
#pragma unroll
	    for (int axis = 0; axis < 3; ++axis)
#pragma unroll
		for (int direction = 0; direction < 2; ++direction)
		{
		    if (!open[axis][direction]) continue;
		    const long long a = c[axis] + (direction == 0 ? -1 : 1);  // cellToCellMap
		    if (a < 0 || a >= g.r[axis]) continue;
		    if (liquidSurface[i + (direction == 0 ? -stride[axis] : stride[axis])] <= 0) liquid = true;
		}
	}
	label = liquid ? MAT_LIQUID : MAT_AIR;
    }
    material[i] = label;
}

// GFS.cpp:717-744 buildValidFaces -> HDK_Utilities.h:137-189 classifyValidFaces, one axis: a face is VALID (1) where its cut-cell
// weight is > 0, both of its cells are in range and at least one of them is LIQUID; INVALID (0) elsewhere (:724)
__global__ void __launch_bounds__(BLOCK) k_fe_valid_faces(float *validFaces, const int32_t *material, const float *cutCell, BaseBox g, int axis)
{
    long long fr[3] = {g.r[0], g.r[1], g.r[2]};
    ++fr[axis];
    const long long f = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (f >= fr[0] * fr[1] * fr[2]) return;
    float v = 0.f;
    if (cutCell[f] > 0)
    {
	const long long c[3] = {f % fr[0], (f / fr[0]) % fr[1], f / (fr[0] * fr[1])};
	if (c[axis] - 1 >= 0 && c[axis] < g.r[axis])
	{
	    const long long fi = c[0] + g.r[0] * (c[1] + g.r[1] * c[2]);
	    const long long stride = axis == 0 ? 1 : (axis == 1 ? g.r[0] : g.r[0] * g.r[1]);
	    if (material[fi - stride] == MAT_LIQUID || material[fi] == MAT_LIQUID) v = 1.f;
	}
    }
    validFaces[f] = v;
}
} // namespace gmg
