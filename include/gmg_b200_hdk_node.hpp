// gmg_b200_hdk_node.hpp -- the steps either side of the solve with the reference's own names and signatures
// over the C ABI of gmg_b200.h, on the caller's own SIM fields:
//
//   HDK::Utilities::buildMaterialCellLabels                  HDK_Utilities.h:49-52   (HDK_Utilities.cpp:87-148)
//   HDK::FreeSurfacePressure::buildValidFaces                 HDK_GeometricFreeSurfacePressureSolver.h:63-66   (.cpp:717-744)
//                            ::buildMGDomainLabels            .h:68-70    (.cpp:746-793)
//                            ::buildMGBoundaryWeights         .h:72-79    (.cpp:796-865)
//                            ::buildRHS                       .h:81-88    (.cpp:868-943)
//                            ::applyOldPressure               .h:90-95    (.cpp:946-997)
//                            ::applySolutionToPressure        .h:97-102   (.cpp:1000-1047)
//                            ::applyPressureGradient          .h:104-111  (.cpp:1050-1131)
//
// In the reference the last seven are private members of the node class; here they are free functions with the same parameter lists,
// so the calls in solveGasSubclass (HDK_GeometricFreeSurfacePressureSolver.cpp:270-413, :640-660) only change their qualification.
// Every function flattens the fields it is given into dense host images, calls libgmg_b200.so (CUDA kernels, csrc/gmg_frontend.cuh)
// and writes the result back through UT_VoxelArray::setValue; there is no CPU arithmetic here and no fallback.
// The one HDK evaluation the reference makes inside these steps stays on the host, made by the HDK itself: solidSurface.getValue(pos) at
// the surface field's cell centres (HDK_Utilities.cpp:21-24) and solidVelocity->getField(axis)->getValue(pos) at the face centres
// (HDK_GeometricFreeSurfacePressureSolver.cpp:918-921) -- skipped where the two fields are aligned (the sample is then the field's own value).
#pragma once

#include <SIM/SIM_RawField.h>
#include <SIM/SIM_RawIndexField.h>
#include <SIM/SIM_VectorField.h>

#include "gmg_b200_hdk.hpp"

namespace GMG_HDK_NAMESPACE
{
namespace B200
{
inline Box wholeBox(const UT_Vector3I &res)
{
    Box b;
    for (int a = 0; a < 3; ++a) b.hi[a] = res[a];
    return b;
}
// dense x-fastest image of a whole voxel array
template <typename T>
inline void image(Dense<T> &out, const UT_VoxelArray<T> &a)
{
    out.reset(a.getVoxelRes());
    flatten(out, a, wholeBox(a.getVoxelRes()));
}
// SIM_RawIndexField (exint) -> the int32 labels of the C ABI
inline void imageLabels(Dense<int32_t> &out, const SIM_RawIndexField &f)
{
    Dense<exint> wide;
    image(wide, *f.field());
    out.reset(f.field()->getVoxelRes());
    const int64_t n = out.count();
    for (int64_t i = 0; i < n; ++i) out.data[i] = int32_t(wide.data[i]);
}
inline void imageLabels(Dense<int32_t> &out, const UT_VoxelArray<int> &a)
{
    Dense<int> d;
    image(d, a);
    out.reset(a.getVoxelRes());
    const int64_t n = out.count();
    for (int64_t i = 0; i < n; ++i) out.data[i] = int32_t(d.data[i]);
}
// every voxel of the dense image back into the voxel array
template <typename T>
inline void store(UT_VoxelArray<T> &a, Dense<T> &in)
{
    unflatten(a, in, wholeBox(a.getVoxelRes()), [](int64_t, int64_t, int64_t) { return true; });
}
// `source` sampled at the sample positions of `at` (the HDK's own interpolation), or its own values where the two are aligned
inline void imageSampledAt(Dense<fpreal32> &out, const SIM_RawField &source, const SIM_RawField &at)
{
    if (source.isAligned(&at)) { image(out, *source.field()); return; }
    const UT_Vector3I res = at.field()->getVoxelRes();
    out.reset(res);
    parallelFor(int(res[2]), [&](int z) {
	for (int y = 0; y < int(res[1]); ++y)
	    for (int x = 0; x < int(res[0]); ++x)
	    {
		UT_Vector3 pos;
		at.indexToPos(x, y, z, pos);
		out.at(x, y, z) = source.getValue(pos);
	    }
    });
}
inline void res3(int64_t out[3], const UT_Vector3I &r) { out[0] = r[0]; out[1] = r[1]; out[2] = r[2]; }
} // namespace B200

// ----------------------------------------------------------------------------------------------------
// HDK::Utilities::buildMaterialCellLabels (HDK_Utilities.h:49-52)
// ----------------------------------------------------------------------------------------------------
namespace Utilities
{
enum FreeSurfaceMaterialLabels { SOLID_CELL, LIQUID_CELL, AIR_CELL };  // HDK_Utilities.h:17

inline void buildMaterialCellLabels(SIM_RawIndexField &materialCellLabels, const SIM_RawField &liquidSurface, const SIM_RawField &solidSurface,
				    const std::array<const SIM_RawField *, 3> &cutCellWeights)
{
    materialCellLabels.match(liquidSurface);  // HDK_Utilities.cpp:98
    B200::Dense<fpreal32> liquid, solid, cut[3];
    B200::image(liquid, *liquidSurface.field());
    B200::imageSampledAt(solid, solidSurface, liquidSurface);
    const float *cp[3];
    for (int a = 0; a < 3; ++a)
    {
	B200::image(cut[a], *cutCellWeights[a]->field());
	cp[a] = cut[a].data;
    }
    int64_t res[3];
    B200::res3(res, liquidSurface.field()->getVoxelRes());
    B200::Dense<int32_t> labels(liquidSurface.field()->getVoxelRes());
    B200::check(gmg_build_material_labels(B200::context(), liquid.data, solid.data, cp, res, labels.data), "gmg_build_material_labels");
    B200::Dense<exint> wide(liquidSurface.field()->getVoxelRes());
    for (int64_t i = 0, n = wide.count(); i < n; ++i) wide.data[i] = labels.data[i];
    B200::store(*materialCellLabels.fieldNC(), wide);
    materialCellLabels.fieldNC()->collapseAllTiles();  // HDK_Utilities.cpp:147
}
} // namespace Utilities

// ----------------------------------------------------------------------------------------------------
// the node's builders (private members of HDK_GeometricFreeSurfacePressureSolver in the reference), same parameter lists
// ----------------------------------------------------------------------------------------------------
namespace FreeSurfacePressure
{
using StoreReal = double;  // HDK_GeometricFreeSurfacePressureSolver.h:18-19
using SolveReal = double;

// .cpp:717-744
inline void buildValidFaces(SIM_VectorField &validFaces, const SIM_RawIndexField &materialCellLabels, const std::array<const SIM_RawField *, 3> &cutCellWeights)
{
    B200::Dense<int32_t> material;
    B200::imageLabels(material, materialCellLabels);
    int64_t res[3];
    B200::res3(res, materialCellLabels.field()->getVoxelRes());
    for (int axis : {0, 1, 2})
    {
	B200::Dense<fpreal32> cut;
	B200::image(cut, *cutCellWeights[axis]->field());
	B200::Dense<fpreal32> valid(cutCellWeights[axis]->field()->getVoxelRes());
	B200::check(gmg_build_valid_faces(B200::context(), material.data, cut.data, res, axis, valid.data), "gmg_build_valid_faces");
	B200::store(*validFaces.getField(axis)->fieldNC(), valid);
	validFaces.getField(axis)->fieldNC()->collapseAllTiles();
    }
}

// .cpp:746-793: mgDomainCellLabels comes in sized like the material labels and filled with EXTERIOR (.cpp:304-309)
inline void buildMGDomainLabels(UT_VoxelArray<int> &mgDomainCellLabels, const SIM_RawIndexField &materialCellLabels)
{
    B200::Dense<int32_t> material;
    B200::imageLabels(material, materialCellLabels);
    int64_t res[3];
    B200::res3(res, materialCellLabels.field()->getVoxelRes());
    B200::Dense<int32_t> labels(materialCellLabels.field()->getVoxelRes());
    B200::check(gmg_build_domain_labels(B200::context(), material.data, res, labels.data), "gmg_build_domain_labels");
    B200::Dense<int> out(materialCellLabels.field()->getVoxelRes());
    for (int64_t i = 0, n = out.count(); i < n; ++i) out.data[i] = int(labels.data[i]);
    B200::store(mgDomainCellLabels, out);
    mgDomainCellLabels.collapseAllTiles();  // .cpp:792
}

// .cpp:796-865, one axis: boundaryWeights comes in sized like the face field and filled with 0 (.cpp:319-322)
inline void buildMGBoundaryWeights(UT_VoxelArray<SolveReal> &boundaryWeights, const SIM_RawField &cutCellWeights, const SIM_RawField &liquidSurface,
				   const SIM_RawField &validFaces, const SIM_RawIndexField & /*materialCellLabels: only asserted*/,
				   const UT_VoxelArray<int> &mgDomainCellLabels, const int axis)
{
    B200::Dense<fpreal32> cut, liquid, valid;
    B200::image(cut, *cutCellWeights.field());
    B200::image(liquid, *liquidSurface.field());
    B200::image(valid, *validFaces.field());
    B200::Dense<int32_t> labels;
    B200::imageLabels(labels, mgDomainCellLabels);
    int64_t res[3];
    B200::res3(res, mgDomainCellLabels.getVoxelRes());
    B200::Dense<SolveReal> w(boundaryWeights.getVoxelRes());
    B200::check(gmg_build_boundary_weights(B200::context(), cut.data, liquid.data, valid.data, labels.data, res, axis, w.data), "gmg_build_boundary_weights");
    // the reference writes the VALID faces only (.cpp:860); everything else keeps the caller's value
    B200::unflatten(boundaryWeights, w, B200::wholeBox(boundaryWeights.getVoxelRes()),
		    [&](int64_t x, int64_t y, int64_t z) { return valid.at(x, y, z) == 1.0f; });
}

// .cpp:868-943: rhsGrid is the EXPANDED grid; only the LIQUID cells are written
inline void buildRHS(UT_VoxelArray<StoreReal> &rhsGrid, const SIM_RawIndexField &materialCellLabels, const SIM_VectorField &velocity,
		     const SIM_VectorField *solidVelocity, const std::array<const SIM_RawField *, 3> &cutCellWeights,
		     const UT_VoxelArray<int> & /*mgDomainCellLabels: only asserted*/, const UT_Vector3I mgExpandedOffset)
{
    B200::Dense<int32_t> material;
    B200::imageLabels(material, materialCellLabels);
    B200::Dense<fpreal32> vel[3], cut[3], solid[3];
    const float *vp[3], *cp[3], *sp[3];
    for (int a = 0; a < 3; ++a)
    {
	B200::image(vel[a], *velocity.getField(a)->field());
	B200::image(cut[a], *cutCellWeights[a]->field());
	vp[a] = vel[a].data;
	cp[a] = cut[a].data;
	sp[a] = nullptr;
	if (solidVelocity)
	{
	    B200::imageSampledAt(solid[a], *solidVelocity->getField(a), *velocity.getField(a));
	    sp[a] = solid[a].data;
	}
    }
    int64_t res[3], expRes[3], off[3];
    B200::res3(res, materialCellLabels.field()->getVoxelRes());
    B200::res3(expRes, rhsGrid.getVoxelRes());
    B200::res3(off, mgExpandedOffset);
    B200::Box box;
    for (int a = 0; a < 3; ++a) { box.lo[a] = off[a]; box.hi[a] = off[a] + res[a]; }
    B200::Dense<StoreReal> grid(rhsGrid.getVoxelRes());
    B200::flatten(grid, rhsGrid, box);
    B200::check(gmg_build_rhs(B200::context(), material.data, vp, cp, solidVelocity ? sp : nullptr, res, expRes, off, grid.data), "gmg_build_rhs");
    B200::unflatten(rhsGrid, grid, box, [&](int64_t x, int64_t y, int64_t z) {
	return material.at(x - off[0], y - off[1], z - off[2]) == Utilities::LIQUID_CELL;
    });
}

// .cpp:946-997
inline void applyOldPressure(UT_VoxelArray<StoreReal> &solutionGrid, const SIM_RawField &pressure, const SIM_RawIndexField &materialCellLabels,
			     const UT_VoxelArray<int> & /*mgDomainCellLabels: only asserted*/, const UT_Vector3I &mgExpandedOffset)
{
    B200::Dense<int32_t> material;
    B200::imageLabels(material, materialCellLabels);
    B200::Dense<fpreal32> pr;
    B200::image(pr, *pressure.field());
    int64_t res[3], expRes[3], off[3];
    B200::res3(res, materialCellLabels.field()->getVoxelRes());
    B200::res3(expRes, solutionGrid.getVoxelRes());
    B200::res3(off, mgExpandedOffset);
    B200::Box box;
    for (int a = 0; a < 3; ++a) { box.lo[a] = off[a]; box.hi[a] = off[a] + res[a]; }
    B200::Dense<StoreReal> grid(solutionGrid.getVoxelRes());
    B200::flatten(grid, solutionGrid, box);
    B200::check(gmg_apply_old_pressure(B200::context(), pr.data, material.data, res, expRes, off, grid.data), "gmg_apply_old_pressure");
    B200::unflatten(solutionGrid, grid, box, [&](int64_t x, int64_t y, int64_t z) {
	return material.at(x - off[0], y - off[1], z - off[2]) == Utilities::LIQUID_CELL;
    });
}

// .cpp:1000-1047
inline void applySolutionToPressure(SIM_RawField &pressure, const SIM_RawIndexField &materialCellLabels, const UT_VoxelArray<int> & /*mgDomainCellLabels*/,
				    const UT_VoxelArray<StoreReal> &solutionGrid, const UT_Vector3I &mgExpandedOffset)
{
    B200::Dense<int32_t> material;
    B200::imageLabels(material, materialCellLabels);
    B200::Dense<fpreal32> pr;
    B200::image(pr, *pressure.field());
    int64_t res[3], expRes[3], off[3];
    B200::res3(res, materialCellLabels.field()->getVoxelRes());
    B200::res3(expRes, solutionGrid.getVoxelRes());
    B200::res3(off, mgExpandedOffset);
    B200::Box box;
    for (int a = 0; a < 3; ++a) { box.lo[a] = off[a]; box.hi[a] = off[a] + res[a]; }
    B200::Dense<StoreReal> grid(solutionGrid.getVoxelRes());
    B200::flatten(grid, solutionGrid, box);
    B200::check(gmg_apply_solution_to_pressure(B200::context(), pr.data, material.data, grid.data, res, expRes, off), "gmg_apply_solution_to_pressure");
    B200::unflatten(*pressure.fieldNC(), pr, B200::wholeBox(pressure.field()->getVoxelRes()),
		    [&](int64_t x, int64_t y, int64_t z) { return material.at(x, y, z) == Utilities::LIQUID_CELL; });
}

// .cpp:1050-1131, one axis: `velocity` is the face field of that axis
inline void applyPressureGradient(SIM_RawField &velocity, const SIM_RawField & /*cutCellWeights: only asserted*/, const SIM_RawField &liquidSurface,
				  const SIM_RawField &pressure, const SIM_RawField &validFaces, const SIM_RawIndexField &materialCellLabels, const int axis)
{
    B200::Dense<int32_t> material;
    B200::imageLabels(material, materialCellLabels);
    B200::Dense<fpreal32> vel, liquid, pr, valid;
    B200::image(vel, *velocity.field());
    B200::image(liquid, *liquidSurface.field());
    B200::image(pr, *pressure.field());
    B200::image(valid, *validFaces.field());
    int64_t res[3];
    B200::res3(res, materialCellLabels.field()->getVoxelRes());
    B200::check(gmg_apply_pressure_gradient(B200::context(), vel.data, liquid.data, pr.data, valid.data, material.data, res, axis), "gmg_apply_pressure_gradient");
    B200::unflatten(*velocity.fieldNC(), vel, B200::wholeBox(velocity.field()->getVoxelRes()),
		    [&](int64_t x, int64_t y, int64_t z) { return valid.at(x, y, z) == 1.0f; });
}
} // namespace FreeSurfacePressure
} // namespace GMG_HDK_NAMESPACE
