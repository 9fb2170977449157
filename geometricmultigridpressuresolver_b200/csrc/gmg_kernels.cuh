// gmg_kernels.cuh -- hand-written sm_100a kernels of the MGPCG hot path.
//
// None of this is a dense contraction, so tensor cores are not used; every kernel here is
// HBM-bound (or latency-bound on coarse levels).  The rules applied: one CTA per 512-cell x 4-plane
// chunk of ACTIVE storage (inactive chunks are never launched), 16-byte vector loads/stores on
// aligned child pairs, labels as bytes, boundary cells handled from a precomputed coefficient
// record list so the full-grid path is a branch-free 7-point stencil, deterministic two-stage
// reductions with warp shuffles.
#pragma once

#include <cooperative_groups.h>

#include "gmg_common.cuh"

namespace gmg
{
namespace cg = cooperative_groups;
// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

__device__ __forceinline__ double warpSum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warpMax(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum in a fixed order; result valid in thread 0.
template <bool IS_MAX = false>
__device__ __forceinline__ double blockReduce(double v)
{
    __shared__ double warpPart[BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = IS_MAX ? warpMax(v) : warpSum(v);
    __syncthreads();
    if (lane == 0) warpPart[warp] = v;
    __syncthreads();
    if (warp == 0)
    {
	v = lane < BLOCK / 32 ? warpPart[lane] : 0.0;
	v = IS_MAX ? warpMax(v) : warpSum(v);
    }
    return v;
}

// Two-stage deterministic grid reduction: every CTA writes its partial, the last CTA to arrive
// (ticket counter) sums the partials in index order and stores the result.  Returns true in the
// finishing CTA's thread 0 only.
template <bool IS_MAX = false>
__device__ __forceinline__ bool gridReduce(double v, double *partials, unsigned *ticket, double *result)
{
    __shared__ bool isLast;
    v = blockReduce<IS_MAX>(v);
    if (threadIdx.x == 0)
    {
	partials[blockIdx.x] = v;
	__threadfence();
	const unsigned t = atomicAdd(ticket, 1u);
	isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return false;
    __threadfence();
    double acc = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += BLOCK)
    {
	const double p = __ldcg(partials + i);
	acc = IS_MAX ? fmax(acc, p) : acc + p;
    }
    acc = blockReduce<IS_MAX>(acc);
    if (threadIdx.x == 0)
    {
	if (result) *result = acc;
	*ticket = 0u;
	return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// device scalars of the PCG loop (one struct in global memory)
// ------------------------------------------------------------------------------------------------
struct Scalars
{
    double rho;      // z.r of the current direction
    double pAp;      // p.Ap
    double rr;       // |r|^2
    double rhoNew;   // z.r after the preconditioner
    double bb;       // |b|^2
    double tmp;      // generic reduction result
    double alpha, beta;
};

// ------------------------------------------------------------------------------------------------
// 7-point operator on the full grid.
//   CTAs [0, nChunks)            : INTERIOR-labelled cells of one 512-cell x 4-plane chunk (branch-free stencil,
//                                   diag 6, all six neighbours active by construction; Ops.h:191-207)
//   CTAs [nChunks, +boundaryCTAs) : BOUNDARY-labelled cells from the coefficient records (Ops.h:208-255)
// Reference arithmetic order kept: lap = -sum over (axis, direction) of c*u, then += diag*u(centre).
// ------------------------------------------------------------------------------------------------
enum StencilMode { SM_JACOBI = 0, SM_APPLY = 1, SM_RESIDUAL = 2 };

struct StencilArgs
{
    const uint8_t *labels;
    const double *in;   // x (Jacobi/residual) or source (apply)
    const double *b;    // rhs (Jacobi/residual)
    double *out;        // Jacobi: new x (out of place); apply: A in; residual: b - A in
    const int32_t *chunks;
    int nChunks;
    int chunksPerPlane;
    int pitch;
    int64_t plane;
    int nz;
    int zlo, zhi;       // z-plane clip [zlo, zhi): planes outside are left untouched (deep-halo sharding)
    // boundary records
    int nBoundary;
    const int32_t *bandIdx;
    const double *bcoef;
    // fused dot(in, A in) for apply
    double *partials;
    unsigned *ticket;
    double *result;
};

template <int MODE>
__device__ __forceinline__ double stencilFinish(double lap, double centre, double rhs, double diag)
{
    if (MODE == SM_APPLY) return lap;
    if (MODE == SM_RESIDUAL) return rhs + (-1.0) * lap;  // addVectors(residual, rhs, residual, -1), Ops.h:731
    double r = rhs - lap;                                 // Ops.h:357-361
    r /= diag;
    return centre + (2.0 / 3.0) * r;
}

// vb = virtual CTA index (blockIdx.x of the stand-alone kernel, a strided index inside the persistent coarse-cycle kernel),
// tid = thread index inside the BLOCK-wide virtual CTA.  Returns this thread's part of dot(in, A in) when DOT.
template <int MODE, bool DOT>
__device__ __forceinline__ double stencilBody(const StencilArgs &a, int vb, int tid)
{
    double acc = 0.0;
    if (vb < a.nChunks)
    {
	const int c = a.chunks[vb];
	const int zb = c / a.chunksPerPlane;
	const int64_t inPlane = int64_t(c - zb * a.chunksPerPlane) * CHUNK_CELLS + 2 * tid;
	if (inPlane < a.plane)
	{
	    const int z0 = zb * CHUNK_Z;
#pragma unroll
	    for (int dz = 0; dz < CHUNK_Z; ++dz)
	    {
		const int z = z0 + dz;
		if (z >= a.zhi) break;
		if (z < a.zlo) continue;
		const int64_t i = int64_t(z) * a.plane + inPlane;
		const uchar2 l = *reinterpret_cast<const uchar2 *>(a.labels + i);
		const bool a0 = (l.x == L_INTERIOR), a1 = (l.y == L_INTERIOR);
		if (!(a0 | a1)) continue;
		const double2 c2 = ld2(a.in + i);
		const double xm = a.in[i - 1], xp = a.in[i + 2];
		const double2 ym = ld2(a.in + i - a.pitch), yp = ld2(a.in + i + a.pitch);
		const double2 zm = ld2(a.in + i - a.plane), zp = ld2(a.in + i + a.plane);
		double2 rhs = make_double2(0.0, 0.0);
		if (MODE != SM_APPLY) rhs = ld2(a.b + i);
		double lap0 = -xm;
		lap0 -= c2.y; lap0 -= ym.x; lap0 -= yp.x; lap0 -= zm.x; lap0 -= zp.x;
		lap0 += 6.0 * c2.x;
		double lap1 = -c2.x;
		lap1 -= xp; lap1 -= ym.y; lap1 -= yp.y; lap1 -= zm.y; lap1 -= zp.y;
		lap1 += 6.0 * c2.y;
		const double o0 = stencilFinish<MODE>(lap0, c2.x, rhs.x, 6.0);
		const double o1 = stencilFinish<MODE>(lap1, c2.y, rhs.y, 6.0);
		if (a0 & a1) st2(a.out + i, make_double2(o0, o1));
		else if (a0) a.out[i] = o0;
		else a.out[i + 1] = o1;
		if (DOT) acc += (a0 ? c2.x * lap0 : 0.0) + (a1 ? c2.y * lap1 : 0.0);
	    }
	}
    }
    else
    {
	const int k = (vb - a.nChunks) * BLOCK + tid;
	const int64_t i = k < a.nBoundary ? int64_t(a.bandIdx[k]) : 0;
	const int z = int(i / a.plane);
	if (k < a.nBoundary && z >= a.zlo && z < a.zhi)
	{
	    const int64_t stride[6] = {-1, 1, -int64_t(a.pitch), int64_t(a.pitch), -a.plane, a.plane};
	    const double centre = a.in[i];
	    double lap = 0.0;
#pragma unroll
	    for (int n = 0; n < 6; ++n)
	    {
		const double cn = a.bcoef[int64_t(n) * a.nBoundary + k];
		if (cn != 0.0) lap -= cn * a.in[i + stride[n]];
	    }
	    const double diag = a.bcoef[int64_t(6) * a.nBoundary + k];
	    lap += diag * centre;
	    const double rhs = (MODE != SM_APPLY) ? a.b[i] : 0.0;
	    a.out[i] = stencilFinish<MODE>(lap, centre, rhs, diag);
	    if (DOT) acc += centre * lap;
	}
    }
    return acc;
}

template <int MODE, bool DOT>
__global__ void __launch_bounds__(BLOCK) k_stencil(const StencilArgs a)
{
    const double acc = stencilBody<MODE, DOT>(a, blockIdx.x, threadIdx.x);
    if (DOT) gridReduce(acc, a.partials, a.ticket, a.result);
}

// ------------------------------------------------------------------------------------------------
// Boundary-band damped Jacobi (Ops.h:524-619).  The reference computes every band cell from the
// current grid into a temporary list, then writes the list back.  Here sweep s reads band
// neighbours from compact array vin (the state after sweep s-1) and frozen non-band neighbours
// from the grid, writing compact array vout -- or the grid itself on the last sweep, which is
// race-free because nobody reads band cells from the grid in that sweep.
// ------------------------------------------------------------------------------------------------
struct BandArgs
{
    double *x;           // grid
    const double *b;     // grid rhs
    const int32_t *bandIdx;
    const int32_t *bandNbr;
    const double *bcoef;
    const double *vin;
    double *vout;
    double *bandB;
    int nBoundary, nBand;
    int pitch;
    int64_t plane;
};

// FROM_COMPACT: centre/band-neighbour values come from vin; TO_GRID: result goes to x[idx];
// FIRST: rhs is gathered from the grid and cached in bandB; ZERO: the grid is known to be all zero.
template <bool FROM_COMPACT, bool TO_GRID, bool FIRST, bool ZERO>
__device__ __forceinline__ void bandBody(const BandArgs &a, int vb, int tid)
{
    const int k = vb * BLOCK + tid;
    if (k >= a.nBand) return;
    const int64_t i = a.bandIdx[k];
    double rhs;
    if (FIRST) { rhs = a.b[i]; a.bandB[k] = rhs; }
    else rhs = a.bandB[k];
    const bool isBoundary = k < a.nBoundary;
    const double diag = isBoundary ? a.bcoef[int64_t(6) * a.nBoundary + k] : 6.0;
    double centre = 0.0, lap = 0.0;
    if (!ZERO)
    {
	centre = FROM_COMPACT ? a.vin[k] : a.x[i];
	const int64_t stride[6] = {-1, 1, -int64_t(a.pitch), int64_t(a.pitch), -a.plane, a.plane};
#pragma unroll
	for (int n = 0; n < 6; ++n)
	{
	    const double cn = isBoundary ? a.bcoef[int64_t(n) * a.nBoundary + k] : 1.0;
	    if (cn != 0.0)
	    {
		int j = -1;
		if (FROM_COMPACT) j = a.bandNbr[int64_t(n) * a.nBand + k];
		const double u = (j >= 0) ? a.vin[j] : a.x[i + stride[n]];
		lap -= cn * u;
	    }
	}
	lap += diag * centre;
    }
    double r = rhs - lap;
    r /= diag;
    const double v = centre + (2.0 / 3.0) * r;
    if (TO_GRID) a.x[i] = v;
    else a.vout[k] = v;
}
template <bool FROM_COMPACT, bool TO_GRID, bool FIRST, bool ZERO>
__global__ void __launch_bounds__(BLOCK) k_band(const BandArgs a)
{
    bandBody<FROM_COMPACT, TO_GRID, FIRST, ZERO>(a, blockIdx.x, threadIdx.x);
}

__global__ void __launch_bounds__(BLOCK) k_band_scatter(double *x, const int32_t *bandIdx, const double *v, int nBand)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k < nBand) x[bandIdx[k]] = v[k];
}

// ------------------------------------------------------------------------------------------------
// Restriction (Ops.h:734-835): coarse (active) = sum_{z,y,x} w[x]w[y]w[z] fine(2c-1+(x,y,z)), weights (1,3,3,1)/8.
// One thread per coarse cell over the chunks holding active coarse cells.
// ------------------------------------------------------------------------------------------------
struct TransferArgs
{
    const uint8_t *fineLabels, *coarseLabels;
    const double *fine;
    const double *coarse;
    double *out;
    const int32_t *chunks;
    int chunksPerPlane;
    int finePitch, coarsePitch;
    int64_t finePlane, coarsePlane;
    int fineNz, coarseNz, coarseNy;
    int shift[3];
    int zlo, zhi;  // clip on the destination's z-planes (coarse planes for restriction, fine planes for prolongation)
};

__device__ __forceinline__ void restrictBody(const TransferArgs &a, int vb, int tid)
{
    const double rw[4] = {1. / 8., 3. / 8., 3. / 8., 1. / 8.};
    const int c = a.chunks[vb];
    const int zb = c / a.chunksPerPlane;
    const int64_t base = int64_t(c - zb * a.chunksPerPlane) * CHUNK_CELLS;
#pragma unroll 1
    for (int dz = 0; dz < CHUNK_Z; ++dz)
    {
	const int cz = zb * CHUNK_Z + dz;
	if (cz >= a.zhi) break;
	if (cz < a.zlo) continue;
#pragma unroll 1
	for (int h = 0; h < 2; ++h)
	{
	    const int64_t inPlane = base + h * BLOCK + tid;
	    if (inPlane >= a.coarsePlane) continue;
	    const int64_t ci = int64_t(cz) * a.coarsePlane + inPlane;
	    const int l = a.coarseLabels[ci];
	    if (!(l == L_INTERIOR || l == L_BOUNDARY)) continue;
	    const int cy = int(inPlane / a.coarsePitch), cx = int(inPlane - int64_t(cy) * a.coarsePitch);
	    const int fx = 2 * (cx - a.shift[0]) - 1, fy = 2 * (cy - a.shift[1]) - 1, fz = 2 * (cz - a.shift[2]) - 1;
	    const double *f = a.fine + (int64_t(fz) * a.finePlane + int64_t(fy) * a.finePitch + fx);
	    double v = 0.0;
#pragma unroll
	    for (int z = 0; z < 4; ++z)
#pragma unroll
		for (int y = 0; y < 4; ++y)
		{
		    const double *row = f + int64_t(z) * a.finePlane + int64_t(y) * a.finePitch;
		    // fx is odd: row[1..2] is an aligned pair
		    const double s0 = row[0];
		    const double2 s12 = ld2(row + 1);
		    const double s3 = row[3];
		    v += rw[0] * rw[y] * rw[z] * s0;
		    v += rw[1] * rw[y] * rw[z] * s12.x;
		    v += rw[2] * rw[y] * rw[z] * s12.y;
		    v += rw[3] * rw[y] * rw[z] * s3;
		}
	    a.out[ci] = v;
	}
    }
}
__global__ void __launch_bounds__(BLOCK) k_restrict(const TransferArgs a) { restrictBody(a, blockIdx.x, threadIdx.x); }

// ------------------------------------------------------------------------------------------------
// Prolongation (Ops.h:873-972): fine (active) += 4 * trilerp(8 coarse cells), fractions .25/.75.
// One thread per aligned fine pair (the two x-children of one coarse cell).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double lerpRef(double v0, double v1, double f) { return (1. - f) * v0 + f * v1; }  // Ops.h:841-848

__device__ __forceinline__ void prolongBody(const TransferArgs &a, int vb, int tid)
{
    const int c = a.chunks[vb];
    const int zb = c / a.chunksPerPlane;
    const int64_t inPlane = int64_t(c - zb * a.chunksPerPlane) * CHUNK_CELLS + 2 * tid;
    if (inPlane >= a.finePlane) return;
    const int fy = int(inPlane / a.finePitch), fx = int(inPlane - int64_t(fy) * a.finePitch);
    const int mx = (fx >> 1) + a.shift[0];
    const int my = (fy >> 1) + a.shift[1];
    const int ys = (fy & 1) ? my : my - 1;
    const double wy = (fy & 1) ? .25 : .75;
#pragma unroll 1
    for (int dz = 0; dz < CHUNK_Z; ++dz)
    {
	const int fz = zb * CHUNK_Z + dz;
	if (fz >= a.zhi) break;
	if (fz < a.zlo) continue;
	const int64_t i = int64_t(fz) * a.finePlane + inPlane;
	const uchar2 l = *reinterpret_cast<const uchar2 *>(a.fineLabels + i);
	const bool a0 = (l.x == L_INTERIOR || l.x == L_BOUNDARY), a1 = (l.y == L_INTERIOR || l.y == L_BOUNDARY);
	if (!(a0 | a1)) continue;
	const int mz = (fz >> 1) + a.shift[2];
	const int zs = (fz & 1) ? mz : mz - 1;
	const double wz = (fz & 1) ? .25 : .75;
	double v[3][2][2];
#pragma unroll
	for (int z = 0; z < 2; ++z)
#pragma unroll
	    for (int y = 0; y < 2; ++y)
	    {
		const double *row = a.coarse + (int64_t(zs + z) * a.coarsePlane + int64_t(ys + y) * a.coarsePitch + mx);
		v[0][y][z] = row[-1];
		v[1][y][z] = row[0];
		v[2][y][z] = row[1];
	    }
	const double2 old = ld2(a.out + i);
	// even child: start = m-1, f = .75; odd child: start = m, f = .25
	const double e = lerpRef(lerpRef(lerpRef(v[0][0][0], v[1][0][0], .75), lerpRef(v[0][1][0], v[1][1][0], .75), wy),
				 lerpRef(lerpRef(v[0][0][1], v[1][0][1], .75), lerpRef(v[0][1][1], v[1][1][1], .75), wy), wz);
	const double o = lerpRef(lerpRef(lerpRef(v[1][0][0], v[2][0][0], .25), lerpRef(v[1][1][0], v[2][1][0], .25), wy),
				 lerpRef(lerpRef(v[1][0][1], v[2][0][1], .25), lerpRef(v[1][1][1], v[2][1][1], .25), wy), wz);
	const double n0 = old.x + 4. * e, n1 = old.y + 4. * o;
	if (a0 & a1) st2(a.out + i, make_double2(n0, n1));
	else if (a0) a.out[i] = n0;
	else a.out[i + 1] = n1;
    }
}
__global__ void __launch_bounds__(BLOCK) k_prolong(const TransferArgs a) { prolongBody(a, blockIdx.x, threadIdx.x); }

// ------------------------------------------------------------------------------------------------
// Coarsest level (MG.cpp:669-692): gather b, x = A^-1 b (dense inverse of the SPD matrix, built on the
// host at setup from an exact Cholesky factor), scatter.  One warp per row, b staged in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_coarse_solve(double *x, const double *b, const int32_t *idx, const double *inv, int n)
{
    extern __shared__ double sb[];
    for (int i = threadIdx.x; i < n; i += BLOCK) sb[i] = b[idx[i]];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    if (row >= n) return;
    const double *r = inv + int64_t(row) * n;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc += r[j] * sb[j];
    acc = warpSum(acc);
    if (lane == 0) x[idx[row]] = acc;
}

// ------------------------------------------------------------------------------------------------
// BLAS-1 over the active chunks.  Vector grids are 0 off the active cells, and every operation
// here maps 0 -> 0, so no label reads are needed.
// ------------------------------------------------------------------------------------------------
struct VecArgs
{
    const int32_t *chunks;
    int chunksPerPlane;
    int64_t plane;
    int nz;
    int zlo, zhi;       // z-plane clip [zlo, zhi)
    double *y;          // destination / first operand
    const double *a;    // second operand
    const double *c;    // third operand
    double *y2;         // second destination (fused CG update)
    double s;           // host scalar
    const Scalars *sc;  // device scalars
    double *partials;
    unsigned *ticket;
    double *result;
};

enum VecOp
{
    VO_AXPY = 0,       // y += s*a
    VO_ADD_SCALED,     // y = a + s*c
    VO_SCALE,          // y *= s
    VO_DOT,            // result = sum y*a
    VO_NORM2,          // result = sum y*y
    VO_MAX,            // result = max(y, 0)
    VO_CG_UPDATE,      // alpha = rho/pAp; y(x) += alpha*a(p); y2(r) -= alpha*c(Ap); result = |r|^2
    VO_CG_DIRECTION,   // beta = rhoNew/rho; y(p) = a(z) + beta*y(p)
    VO_COPY            // y = a
};

template <int OP>
__global__ void __launch_bounds__(BLOCK) k_vec(const VecArgs v)
{
    const int c = v.chunks[blockIdx.x];
    const int zb = c / v.chunksPerPlane;
    const int64_t inPlane = int64_t(c - zb * v.chunksPerPlane) * CHUNK_CELLS + 2 * threadIdx.x;
    double acc = 0.0;
    double s = v.s;
    if (OP == VO_CG_UPDATE) s = v.sc->rho / v.sc->pAp;
    if (OP == VO_CG_DIRECTION) s = v.sc->rhoNew / v.sc->rho;
    if (inPlane < v.plane)
    {
#pragma unroll
	for (int dz = 0; dz < CHUNK_Z; ++dz)
	{
	    const int z = zb * CHUNK_Z + dz;
	    if (z >= v.zhi) break;
	    if (z < v.zlo) continue;
	    const int64_t i = int64_t(z) * v.plane + inPlane;
	    if (OP == VO_AXPY)
	    {
		double2 y = ld2(v.y + i);
		const double2 a = ld2(v.a + i);
		y.x = y.x + s * a.x; y.y = y.y + s * a.y;
		st2(v.y + i, y);
	    }
	    else if (OP == VO_ADD_SCALED)
	    {
		const double2 a = ld2(v.a + i), cc = ld2(v.c + i);
		st2(v.y + i, make_double2(a.x + s * cc.x, a.y + s * cc.y));
	    }
	    else if (OP == VO_SCALE)
	    {
		double2 y = ld2(v.y + i);
		st2(v.y + i, make_double2(s * y.x, s * y.y));
	    }
	    else if (OP == VO_DOT)
	    {
		const double2 y = ld2(v.y + i), a = ld2(v.a + i);
		acc += y.x * a.x; acc += y.y * a.y;
	    }
	    else if (OP == VO_NORM2)
	    {
		const double2 y = ld2(v.y + i);
		acc += y.x * y.x; acc += y.y * y.y;
	    }
	    else if (OP == VO_MAX)
	    {
		const double2 y = ld2(v.y + i);
		acc = fmax(acc, fmax(y.x, y.y));
	    }
	    else if (OP == VO_CG_UPDATE)
	    {
		double2 x = ld2(v.y + i), r = ld2(v.y2 + i);
		const double2 p = ld2(v.a + i), t = ld2(v.c + i);
		x.x = x.x + s * p.x; x.y = x.y + s * p.y;
		r.x = r.x + (-s) * t.x; r.y = r.y + (-s) * t.y;
		st2(v.y + i, x);
		st2(v.y2 + i, r);
		acc += r.x * r.x; acc += r.y * r.y;
	    }
	    else if (OP == VO_CG_DIRECTION)
	    {
		const double2 zz = ld2(v.a + i), p = ld2(v.y + i);
		st2(v.y + i, make_double2(zz.x + s * p.x, zz.y + s * p.y));
	    }
	    else if (OP == VO_COPY)
	    {
		st2(v.y + i, ld2(v.a + i));
	    }
	}
    }
    if (OP == VO_DOT || OP == VO_NORM2 || OP == VO_CG_UPDATE) gridReduce<false>(acc, v.partials, v.ticket, v.result);
    if (OP == VO_MAX) gridReduce<true>(acc, v.partials, v.ticket, v.result);
}

// zero the active chunks of a grid (x = 0 at the start of a V-cycle level, MG.cpp:439-440, :566)
__device__ __forceinline__ void zeroBody(double *y, const int32_t *chunks, int chunksPerPlane, int64_t plane, int nz, int vb, int tid)
{
    const int c = chunks[vb];
    const int zb = c / chunksPerPlane;
    const int64_t inPlane = int64_t(c - zb * chunksPerPlane) * CHUNK_CELLS + 2 * tid;
    if (inPlane >= plane) return;
#pragma unroll
    for (int dz = 0; dz < CHUNK_Z; ++dz)
    {
	const int z = zb * CHUNK_Z + dz;
	if (z >= nz) break;
	st2(y + int64_t(z) * plane + inPlane, make_double2(0.0, 0.0));
    }
}
__global__ void __launch_bounds__(BLOCK) k_zero(double *y, const int32_t *chunks, int chunksPerPlane, int64_t plane, int nz)
{
    zeroBody(y, chunks, chunksPerPlane, plane, nz, blockIdx.x, threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// Persistent coarse sub-V-cycle.  Below a few hundred thousand cells a level is pure launch latency
// (19 dependent launches of ~3 us each), so levels [first, last] -- down-stroke, direct solve, up-stroke --
// run inside ONE kernel: a single thread-block cluster whose CTAs walk the same virtual-CTA bodies as
// the stand-alone kernels and meet at the hardware cluster barrier (release/acquire at cluster scope,
// which orders the global-memory traffic between the steps) instead of at kernel boundaries.
// Same arithmetic, same order as the per-kernel path: results are bitwise identical.
// ------------------------------------------------------------------------------------------------
constexpr int CYCLE_THREADS = 1024;

struct DevLevel
{
    const uint8_t *labels;
    double *x, *xAlt, *b, *r;
    const int32_t *chunksInterior, *chunksActive;
    int nChunksInterior, nChunksActive;
    int chunksPerPlane, pitch, ny, nz;
    int64_t plane;
    int nBoundary, nBand;
    const int32_t *bandIdx, *bandNbr;
    const double *bcoef;
    double *bandV0, *bandV1, *bandB;
    int shift[3];
};

struct CycleArgs
{
    const DevLevel *lv;  // device array indexed by level
    int first, last;     // `last` is the direct-solve level
    int bandSweeps;
    const int32_t *coarseIdx;
    const double *coarseInv;
    int nCoarse;
};

template <typename F>
__device__ __forceinline__ void forVirtualCtas(int nvb, const F &f)
{
    constexpr int PER = CYCLE_THREADS / BLOCK;
    const int sub = threadIdx.x / BLOCK, tid = threadIdx.x % BLOCK;
    for (int vb = blockIdx.x * PER + sub; vb < nvb; vb += gridDim.x * PER) f(vb, tid);
}

__device__ __forceinline__ void cycleSync() { cg::this_cluster().sync(); }

__device__ __forceinline__ StencilArgs devStencilArgs(const DevLevel &L, const double *in, const double *b, double *out)
{
    StencilArgs a;
    a.labels = L.labels; a.in = in; a.b = b; a.out = out;
    a.chunks = L.chunksInterior; a.nChunks = L.nChunksInterior; a.chunksPerPlane = L.chunksPerPlane;
    a.pitch = L.pitch; a.plane = L.plane; a.nz = L.nz; a.zlo = 0; a.zhi = L.nz;
    a.nBoundary = L.nBoundary; a.bandIdx = L.bandIdx; a.bcoef = L.bcoef;
    a.partials = nullptr; a.ticket = nullptr; a.result = nullptr;
    return a;
}

// `sweeps` band sweeps on grid x, ending on a cluster barrier (mirrors launchBand on the host)
__device__ __forceinline__ void cycleBand(const DevLevel &L, double *x, const double *b, int sweeps, bool zeroGrid)
{
    if (L.nBand == 0 || sweeps <= 0) return;
    BandArgs a;
    a.x = x; a.b = b; a.bandIdx = L.bandIdx; a.bandNbr = L.bandNbr; a.bcoef = L.bcoef; a.bandB = L.bandB;
    a.nBoundary = L.nBoundary; a.nBand = L.nBand; a.pitch = L.pitch; a.plane = L.plane;
    const int nvb = (L.nBand + BLOCK - 1) / BLOCK;
    double *cur = L.bandV0, *nxt = L.bandV1;
    a.vin = nullptr;
    a.vout = cur;
    if (zeroGrid) forVirtualCtas(nvb, [&](int vb, int tid) { bandBody<false, false, true, true>(a, vb, tid); });
    else forVirtualCtas(nvb, [&](int vb, int tid) { bandBody<false, false, true, false>(a, vb, tid); });
    cycleSync();
    if (sweeps == 1)
    {
	forVirtualCtas(nvb, [&](int vb, int tid) { const int k = vb * BLOCK + tid; if (k < L.nBand) x[L.bandIdx[k]] = cur[k]; });
	cycleSync();
	return;
    }
    for (int sw = 2; sw <= sweeps; ++sw)
    {
	a.vin = cur;
	a.vout = nxt;
	if (sw == sweeps) forVirtualCtas(nvb, [&](int vb, int tid) { bandBody<true, true, false, false>(a, vb, tid); });
	else forVirtualCtas(nvb, [&](int vb, int tid) { bandBody<true, false, false, false>(a, vb, tid); });
	cycleSync();
	double *t = cur; cur = nxt; nxt = t;
    }
}

// band sweeps, interior Jacobi cur -> alt, band sweeps on alt (MG.cpp:557-667 / :695-784); the result is in alt
__device__ __forceinline__ void cycleSmooth(const DevLevel &L, double *cur, double *alt, int sweeps, bool zeroGrid)
{
    cycleBand(L, cur, L.b, sweeps, zeroGrid);
    const StencilArgs a = devStencilArgs(L, cur, L.b, alt);
    forVirtualCtas(L.nChunksInterior + (L.nBoundary + BLOCK - 1) / BLOCK, [&](int vb, int tid) { stencilBody<SM_JACOBI, false>(a, vb, tid); });
    cycleSync();
    cycleBand(L, alt, L.b, sweeps, false);
}

__device__ __forceinline__ TransferArgs devTransferArgs(const DevLevel &F, const DevLevel &C)
{
    TransferArgs a;
    a.fineLabels = F.labels; a.coarseLabels = C.labels;
    a.finePitch = F.pitch; a.coarsePitch = C.pitch; a.finePlane = F.plane; a.coarsePlane = C.plane;
    a.fineNz = F.nz; a.coarseNz = C.nz; a.coarseNy = C.ny;
    a.shift[0] = F.shift[0]; a.shift[1] = F.shift[1]; a.shift[2] = F.shift[2];
    a.fine = nullptr; a.coarse = nullptr; a.out = nullptr; a.chunks = nullptr; a.chunksPerPlane = 0; a.zlo = 0; a.zhi = 0;
    return a;
}

__global__ void __launch_bounds__(CYCLE_THREADS, 1) k_coarse_cycle(const CycleArgs c)
{
    extern __shared__ double sb[];
    // ---- down-stroke
    for (int l = c.first; l < c.last; ++l)
    {
	const DevLevel &L = c.lv[l];
	const DevLevel &C = c.lv[l + 1];
	forVirtualCtas(L.nChunksActive, [&](int vb, int tid) { zeroBody(L.x, L.chunksActive, L.chunksPerPlane, L.plane, L.nz, vb, tid); });
	cycleSync();
	cycleSmooth(L, L.x, L.xAlt, c.bandSweeps, true);
	{
	    const StencilArgs a = devStencilArgs(L, L.xAlt, L.b, L.r);
	    forVirtualCtas(L.nChunksInterior + (L.nBoundary + BLOCK - 1) / BLOCK, [&](int vb, int tid) { stencilBody<SM_RESIDUAL, false>(a, vb, tid); });
	}
	cycleSync();
	{
	    TransferArgs a = devTransferArgs(L, C);
	    a.fine = L.r; a.out = C.b; a.chunks = C.chunksActive; a.chunksPerPlane = C.chunksPerPlane; a.zlo = 0; a.zhi = C.nz;
	    forVirtualCtas(C.nChunksActive, [&](int vb, int tid) { restrictBody(a, vb, tid); });
	}
	cycleSync();
    }
    // ---- direct solve on the coarsest level: x = A^-1 b
    {
	const DevLevel &L = c.lv[c.last];
	const int n = c.nCoarse;
	for (int i = threadIdx.x; i < n; i += CYCLE_THREADS) sb[i] = L.b[c.coarseIdx[i]];
	__syncthreads();
	const int lane = threadIdx.x & 31;
	for (int row = blockIdx.x * (CYCLE_THREADS / 32) + (threadIdx.x >> 5); row < n; row += gridDim.x * (CYCLE_THREADS / 32))
	{
	    const double *r = c.coarseInv + int64_t(row) * n;
	    double acc = 0.0;
	    for (int j = lane; j < n; j += 32) acc += r[j] * sb[j];
	    acc = warpSum(acc);
	    if (lane == 0) L.x[c.coarseIdx[row]] = acc;
	}
	cycleSync();
    }
    // ---- up-stroke: x_l (in xAlt after the down-stroke) += P x_{l+1}; smooth back into x
    for (int l = c.last - 1; l >= c.first; --l)
    {
	const DevLevel &L = c.lv[l];
	const DevLevel &C = c.lv[l + 1];
	{
	    TransferArgs a = devTransferArgs(L, C);
	    a.coarse = C.x; a.out = L.xAlt; a.chunks = L.chunksActive; a.chunksPerPlane = L.chunksPerPlane; a.zlo = 0; a.zhi = L.nz;
	    forVirtualCtas(L.nChunksActive, [&](int vb, int tid) { prolongBody(a, vb, tid); });
	}
	cycleSync();
	cycleSmooth(L, L.xAlt, L.x, c.bandSweeps, false);
    }
}

// ------------------------------------------------------------------------------------------------
// Setup kernels: labels (bit-exact integer work)
// ------------------------------------------------------------------------------------------------
struct BoxArgs
{
    int n[3];
    int pitch;
    int64_t plane, total;
    int org[3];
    int64_t res[3];
};

// staging int32 (dense box, x-fastest, row pitch = n[0]) -> byte labels with EXTERIOR outside [validLo, validHi)
__global__ void __launch_bounds__(BLOCK) k_labels_from_i32(uint8_t *labels, const int32_t *staging, BoxArgs g, int lo0, int lo1, int lo2,
							  int hi0, int hi1, int hi2)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= g.total) return;
    const int z = int(i / g.plane);
    const int64_t rem = i - int64_t(z) * g.plane;
    const int y = int(rem / g.pitch), x = int(rem - int64_t(y) * g.pitch);
    int l = L_EXTERIOR;
    if (x >= lo0 && x < hi0 && y >= lo1 && y < hi1 && z >= lo2 && z < hi2)
	l = staging[(int64_t(z - lo2) * (hi1 - lo1) + (y - lo1)) * (hi0 - lo0) + (x - lo0)];
    labels[i] = uint8_t(l);
}

__global__ void __launch_bounds__(BLOCK) k_labels_to_i32(int32_t *staging, const uint8_t *labels, BoxArgs g, int lo0, int lo1, int lo2, int hi0,
							int hi1, int hi2)
{
    const int64_t nx = hi0 - lo0, ny = hi1 - lo1, nz = hi2 - lo2;
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= nx * ny * nz) return;
    const int z = int(i / (nx * ny));
    const int64_t rem = i - int64_t(z) * nx * ny;
    const int y = int(rem / nx), x = int(rem - int64_t(y) * nx);
    staging[i] = labels[int64_t(z + lo2) * g.plane + int64_t(y + lo1) * g.pitch + (x + lo0)];
}

// same for doubles: dense staging box <-> pitched storage (zero outside)
// maskLabels (nullable): force 0 on non-active cells, which restores the vector-grid invariant for host input
__global__ void __launch_bounds__(BLOCK) k_values_from_staging(double *dst, const double *staging, const uint8_t *maskLabels, BoxArgs g, int lo0,
							      int lo1, int lo2, int hi0, int hi1, int hi2)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= g.total) return;
    const int z = int(i / g.plane);
    const int64_t rem = i - int64_t(z) * g.plane;
    const int y = int(rem / g.pitch), x = int(rem - int64_t(y) * g.pitch);
    double v = 0.0;
    if (x >= lo0 && x < hi0 && y >= lo1 && y < hi1 && z >= lo2 && z < hi2)
	v = staging[(int64_t(z - lo2) * (hi1 - lo1) + (y - lo1)) * (hi0 - lo0) + (x - lo0)];
    if (maskLabels)
    {
	const int l = maskLabels[i];
	if (!(l == L_INTERIOR || l == L_BOUNDARY)) v = 0.0;
    }
    dst[i] = v;
}

__global__ void __launch_bounds__(BLOCK) k_values_to_staging(double *staging, const double *src, BoxArgs g, int lo0, int lo1, int lo2, int hi0,
							    int hi1, int hi2)
{
    const int64_t nx = hi0 - lo0, ny = hi1 - lo1, nz = hi2 - lo2;
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= nx * ny * nz) return;
    const int z = int(i / (nx * ny));
    const int64_t rem = i - int64_t(z) * nx * ny;
    const int y = int(rem / nx), x = int(rem - int64_t(y) * nx);
    staging[i] = src[int64_t(z + lo2) * g.plane + int64_t(y + lo1) * g.pitch + (x + lo0)];
}

// base labels -> expanded-box labels (Ops.h:1404-1453): non-exterior copied as INTERIOR / DIRICHLET
__global__ void __launch_bounds__(BLOCK) k_expand_labels(int32_t *out, const int32_t *base, int64_t n)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= n) return;
    const int l = base[i];
    out[i] = (l == L_EXTERIOR) ? L_EXTERIOR : (l == L_INTERIOR ? L_INTERIOR : L_DIRICHLET);
}
// base weights -> expanded weights (Ops.h:1488-1571): copy weights > 0, else 0
__global__ void __launch_bounds__(BLOCK) k_expand_weights(double *out, const double *base, int64_t n)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= n) return;
    const double w = base[i];
    out[i] = w > 0 ? w : 0.0;
}

__device__ __forceinline__ int labelAt(const uint8_t *labels, const BoxArgs &g, int x, int y, int z)
{
    if (x < 0 || y < 0 || z < 0 || x >= g.n[0] || y >= g.n[1] || z >= g.n[2]) return L_EXTERIOR;
    return labels[int64_t(z) * g.plane + int64_t(y) * g.pitch + x];
}

// setBoundaryCellLabels (Ops.h:1574-1644): INTERIOR -> BOUNDARY next to DIRICHLET/EXTERIOR or a face weight != 1.
// w[a] is stored per cell = weight of the cell's BACKWARD face along a; the forward face is the next cell's.
// Out of place (in != out), so the result does not depend on evaluation order.
__global__ void __launch_bounds__(BLOCK) k_set_boundary(uint8_t *out, const uint8_t *in, const double *w0, const double *w1, const double *w2,
						       BoxArgs g)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= g.total) return;
    int l = in[i];
    if (l == L_INTERIOR)
    {
	const int z = int(i / g.plane);
	const int64_t rem = i - int64_t(z) * g.plane;
	const int y = int(rem / g.pitch), x = int(rem - int64_t(y) * g.pitch);
	bool isB = false;
	const int d[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
	const int64_t stride[3] = {1, g.pitch, g.plane};
	const double *w[3] = {w0, w1, w2};
#pragma unroll
	for (int n = 0; n < 6; ++n)
	{
	    const int nl = labelAt(in, g, x + d[n][0], y + d[n][1], z + d[n][2]);
	    if (nl == L_DIRICHLET || nl == L_EXTERIOR) isB = true;
	    else
	    {
		const int axis = n >> 1;
		const double wt = (n & 1) ? w[axis][i + stride[axis]] : w[axis][i];
		if (wt != 1) isB = true;
	    }
	}
	if (isB) l = L_BOUNDARY;
    }
    out[i] = uint8_t(l);
}

// buildCoarseCellLabels pass 1 (Ops.cpp:42-105)
__global__ void __launch_bounds__(BLOCK) k_coarsen1(uint8_t *coarse, const uint8_t *fine, BoxArgs cg, BoxArgs fg, int s0, int s1, int s2)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= cg.total) return;
    const int z = int(i / cg.plane);
    const int64_t rem = i - int64_t(z) * cg.plane;
    const int y = int(rem / cg.pitch), x = int(rem - int64_t(y) * cg.pitch);
    int l = L_EXTERIOR;
    if (x < cg.n[0])
    {
	const int fx = 2 * (x - s0), fy = 2 * (y - s1), fz = 2 * (z - s2);
	bool hasD = false, hasI = false;
#pragma unroll
	for (int k = 0; k < 8; ++k)
	{
	    const int fl = labelAt(fine, fg, fx + (k & 1), fy + ((k >> 1) & 1), fz + (k >> 2));
	    if (fl == L_DIRICHLET) hasD = true;
	    else if (fl == L_INTERIOR || fl == L_BOUNDARY) hasI = true;
	}
	l = hasD ? L_DIRICHLET : (hasI ? L_INTERIOR : L_EXTERIOR);
    }
    coarse[i] = uint8_t(l);
}
// pass 2 (Ops.cpp:107-158), out of place
__global__ void __launch_bounds__(BLOCK) k_coarsen2(uint8_t *out, const uint8_t *in, BoxArgs g)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= g.total) return;
    int l = in[i];
    if (l == L_INTERIOR)
    {
	const int z = int(i / g.plane);
	const int64_t rem = i - int64_t(z) * g.plane;
	const int y = int(rem / g.pitch), x = int(rem - int64_t(y) * g.pitch);
	const int d[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
	bool isB = false;
#pragma unroll
	for (int n = 0; n < 6; ++n)
	{
	    const int nl = labelAt(in, g, x + d[n][0], y + d[n][1], z + d[n][2]);
	    if (nl == L_EXTERIOR || nl == L_DIRICHLET) isB = true;
	}
	if (isB) l = L_BOUNDARY;
    }
    out[i] = uint8_t(l);
}

// ------------------------------------------------------------------------------------------------
// Setup kernels: boundary band (Ops.cpp:165-469) and coefficient records
// ------------------------------------------------------------------------------------------------
// mask bit0 = visited. layer 0: BOUNDARY cells.
__global__ void __launch_bounds__(BLOCK) k_band_init(uint8_t *mask, const uint8_t *labels, int64_t total)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i < total) mask[i] = (labels[i] == L_BOUNDARY) ? 1 : 0;
}
// next layer: unvisited INTERIOR cells with a visited 6-neighbour
__global__ void __launch_bounds__(BLOCK) k_band_dilate(uint8_t *out, const uint8_t *in, const uint8_t *labels, BoxArgs g)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= g.total) return;
    uint8_t m = in[i];
    if (!m && labels[i] == L_INTERIOR)
    {
	// an INTERIOR cell is never on the storage border, so all six neighbours are in range
	if (in[i - 1] | in[i + 1] | in[i - g.pitch] | in[i + g.pitch] | in[i - g.plane] | in[i + g.plane]) m = 1;
    }
    out[i] = m;
}
// flags for the two compactions: which = 0 -> BOUNDARY cells, 1 -> INTERIOR cells of the band
__global__ void __launch_bounds__(BLOCK) k_band_flags(uint8_t *flags, const uint8_t *mask, const uint8_t *labels, int which, int64_t total)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= total) return;
    const bool isB = labels[i] == L_BOUNDARY;
    flags[i] = (mask[i] && (which == 0 ? isB : !isB)) ? 1 : 0;
}
__global__ void __launch_bounds__(BLOCK) k_fill_i32(int32_t *p, int32_t v, int64_t n)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(BLOCK) k_band_pos(int32_t *pos, const int32_t *bandIdx, int nBand)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k < nBand) pos[bandIdx[k]] = k;
}
__global__ void __launch_bounds__(BLOCK) k_band_nbr(int32_t *bandNbr, const int32_t *pos, const int32_t *bandIdx, int nBand, int pitch, int64_t plane)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nBand) return;
    const int64_t i = bandIdx[k];
    const int64_t stride[6] = {-1, 1, -int64_t(pitch), int64_t(pitch), -plane, plane};
#pragma unroll
    for (int n = 0; n < 6; ++n) bandNbr[int64_t(n) * nBand + k] = pos[i + stride[n]];
}
// coefficient record of a BOUNDARY cell (Ops.h:208-255): c_n = 1 (INTERIOR nbr), w (BOUNDARY nbr), 0 otherwise;
// diag = sum of 1 (INTERIOR), w (BOUNDARY), w (DIRICHLET).  w == nullptr (coarse levels) means weight 1.
__global__ void __launch_bounds__(BLOCK) k_band_coef(double *bcoef, const int32_t *bandIdx, int nBoundary, const uint8_t *labels,
						    const double *w0, const double *w1, const double *w2, int pitch, int64_t plane)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nBoundary) return;
    const int64_t i = bandIdx[k];
    const int64_t stride[3] = {1, pitch, plane};
    const double *w[3] = {w0, w1, w2};
    double diag = 0.0;
#pragma unroll
    for (int n = 0; n < 6; ++n)
    {
	const int axis = n >> 1;
	const int64_t j = (n & 1) ? i + stride[axis] : i - stride[axis];
	const int nl = labels[j];
	double wt = 1.0;
	if (w0) wt = (n & 1) ? w[axis][j] : w[axis][i];
	double cn = 0.0;
	if (nl == L_INTERIOR) { cn = 1.0; diag += 1.0; }
	else if (nl == L_BOUNDARY) { cn = wt; diag += wt; }
	else if (nl == L_DIRICHLET) { diag += wt; }
	bcoef[int64_t(n) * nBoundary + k] = cn;
    }
    bcoef[int64_t(6) * nBoundary + k] = diag;
}

// chunk flags: bit0 = chunk holds an INTERIOR cell, bit1 = chunk holds an active cell
__global__ void __launch_bounds__(BLOCK) k_chunk_flags(uint8_t *flagInterior, uint8_t *flagActive, const uint8_t *labels, int chunksPerPlane,
						      int64_t plane, int nz)
{
    const int c = blockIdx.x;
    const int zb = c / chunksPerPlane;
    const int64_t base = int64_t(c - zb * chunksPerPlane) * CHUNK_CELLS;
    int fi = 0, fa = 0;
    for (int dz = 0; dz < CHUNK_Z; ++dz)
    {
	const int z = zb * CHUNK_Z + dz;
	if (z >= nz) break;
	for (int h = 0; h < 2; ++h)
	{
	    const int64_t inPlane = base + h * BLOCK + threadIdx.x;
	    if (inPlane >= plane) continue;
	    const int l = labels[int64_t(z) * plane + inPlane];
	    fi |= (l == L_INTERIOR);
	    fa |= (l == L_INTERIOR || l == L_BOUNDARY);
	}
    }
    fi = __syncthreads_or(fi);
    fa = __syncthreads_or(fa);
    if (threadIdx.x == 0) { flagInterior[c] = uint8_t(fi); flagActive[c] = uint8_t(fa); }
}

// counts of INTERIOR and active cells (for roofline accounting)
__global__ void __launch_bounds__(BLOCK) k_count_labels(const uint8_t *labels, int64_t total, unsigned long long *counts)
{
    unsigned long long ci = 0, cb = 0;
    for (int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x; i < total; i += int64_t(gridDim.x) * BLOCK)
    {
	const int l = labels[i];
	ci += (l == L_INTERIOR);
	cb += (l == L_BOUNDARY);
    }
    ci = __reduce_add_sync(0xffffffffu, unsigned(ci));
    cb = __reduce_add_sync(0xffffffffu, unsigned(cb));
    if ((threadIdx.x & 31) == 0)
    {
	atomicAdd(&counts[0], ci);
	atomicAdd(&counts[1], cb);
    }
}

// sort key of the reference's boundary list (Ops.cpp:441-466): (16^3-tile linear index in the EXPANDED grid, z, y, x)
__global__ void __launch_bounds__(BLOCK) k_band_keys(unsigned long long *keys, const int32_t *bandIdx, int nBand, BoxArgs g)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nBand) return;
    const int64_t i = bandIdx[k];
    const int z = int(i / g.plane);
    const int64_t rem = i - int64_t(z) * g.plane;
    const int y = int(rem / g.pitch), x = int(rem - int64_t(y) * g.pitch);
    const unsigned long long ex = x + g.org[0], ey = y + g.org[1], ez = z + g.org[2];
    const unsigned long long tilesX = (g.res[0] + 15) >> 4, tilesY = (g.res[1] + 15) >> 4;
    const unsigned long long tile = ((ez >> 4) * tilesY + (ey >> 4)) * tilesX + (ex >> 4);
    keys[k] = (tile << 36) | (ez << 24) | (ey << 12) | ex;
}

} // namespace gmg
