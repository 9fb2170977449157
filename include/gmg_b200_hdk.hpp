// gmg_b200_hdk.hpp -- C++ host facade with the reference's own names and signatures over the C ABI
// of gmg_b200.h.  This is what a maintainer of rgoldade/GeometricMultigridPressureSolver includes
// INSTEAD of HDK_GeometricMultigridOperators.h / HDK_GeometricMultigridPoissonSolver.h /
// HDK_GeometricCGPoissonSolver.h; the callers in
// HDK_GeometricFreeSurfacePressureSolver.cpp:344-484 and HDK_TestGeometricMultigrid.cpp:170-204,
// :746-832 compile unchanged against it.
//
//   * containers are the caller's: UT_VoxelArray / UT_Array / UT_Vector3I from the Houdini HDK
//     (<UT/UT_VoxelArray.h>); only the members the reference itself uses are touched;
//   * every function here marshals tiles <-> dense host buffers and calls libgmg_b200.so; there is
//     no CPU arithmetic in this header and no fallback: without the library or a CUDA device every
//     entry point throws HDK::B200::Error;
//   * fp64 only (the reference fixes StoreReal = SolveReal = double, MG.h:14-15, GFS.h:18-19).
//
// Fast path: GeometricMultigridPoissonSolver keeps the hierarchy on the device, and
// solveGeometricConjugateGradient(solver, x, b, tol, maxIt) runs the whole PCG on the device with one
// upload of (x0, b) and one download of x.  The stateless operator functions (applyPoissonMatrix,
// jacobiPoissonSmoother, ...) exist for source compatibility: each call builds a device domain from the
// label grid it is given, so they cost a label upload per call -- hold a
// GeometricMultigridOperators::DeviceDomain to amortise that.
//
// The namespace is HDK by default; define GMG_HDK_NAMESPACE to something else to compile this facade
// next to the reference's own headers (tests/cpp/test_facade.cpp does, for parity checks).
#pragma once

#include <UT/UT_VoxelArray.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "gmg_b200.h"

#ifndef GMG_HDK_NAMESPACE
#define GMG_HDK_NAMESPACE HDK
#endif

namespace GMG_HDK_NAMESPACE
{
// ----------------------------------------------------------------------------------------------------
// runtime: context, errors, tile <-> dense marshalling
// ----------------------------------------------------------------------------------------------------
namespace B200
{
struct Error : std::runtime_error
{
    int status;
    Error(int st, const std::string &what) : std::runtime_error(what), status(st) {}
};

inline void check(int st, const char *where)
{
    if (st != GMG_OK) throw Error(st, std::string(where) + ": " + gmg_last_error());
}

// one context per process and device (the reference is called from one Houdini cook thread)
inline int &deviceOrdinal()
{
    static int d = [] { const char *e = std::getenv("GMG_DEVICE"); return e ? std::atoi(e) : 0; }();
    return d;
}
inline bool &verbose()
{
    static bool v = [] { const char *e = std::getenv("GMG_VERBOSE"); return e && e[0] == '1'; }();
    return v;
}
inline gmg_ctx *context()
{
    struct Holder
    {
	gmg_ctx *ctx = nullptr;
	Holder() { check(gmg_ctx_create(deviceOrdinal(), nullptr, &ctx), "gmg_ctx_create"); }
	~Holder() { gmg_ctx_destroy(ctx); }
    };
    static Holder h;
    return h.ctx;
}

struct Box
{
    int64_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    bool empty() const { return hi[0] <= lo[0] || hi[1] <= lo[1] || hi[2] <= lo[2]; }
};

template <typename Fn>
inline void parallelFor(int n, const Fn &fn)
{
    const int nt = std::max(1, std::min<int>(n, int(std::min(16u, std::max(1u, std::thread::hardware_concurrency())))));
    if (nt == 1) { for (int i = 0; i < n; ++i) fn(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
	th.emplace_back([&fn, t, nt, n]() { for (int i = t; i < n; i += nt) fn(i); });
    for (auto &t : th) t.join();
}

// Dense x-fastest host image of a voxel array; pages that are never written stay untouched zero pages.
template <typename T>
struct Dense
{
    int64_t res[3] = {0, 0, 0};
    T *data = nullptr;
    Dense() {}
    explicit Dense(const UT_Vector3I &r) { reset(r); }
    Dense(const Dense &) = delete;
    Dense &operator=(const Dense &) = delete;
    ~Dense() { std::free(data); }
    void reset(const UT_Vector3I &r)
    {
	std::free(data);
	for (int a = 0; a < 3; ++a) res[a] = r[a];
	data = static_cast<T *>(std::calloc(size_t(res[0]) * res[1] * res[2], sizeof(T)));
	if (!data) throw Error(GMG_ERR_INVALID, "out of host memory for a dense grid image");
    }
    int64_t count() const { return res[0] * res[1] * res[2]; }
    T &at(int64_t x, int64_t y, int64_t z) { return data[x + res[0] * (y + res[1] * z)]; }
};

// visits every 16^3 tile overlapping [lo,hi): fn(tile, x0, y0, z0) with the tile's voxel origin
template <typename T, typename Fn>
inline void forTilesInBox(const UT_VoxelArray<T> &a, const Box &b, const Fn &fn)
{
    const int t0[3] = {int(b.lo[0] >> 4), int(b.lo[1] >> 4), int(b.lo[2] >> 4)};
    const int t1[3] = {int((b.hi[0] + 15) >> 4), int((b.hi[1] + 15) >> 4), int((b.hi[2] + 15) >> 4)};
    const int nz = t1[2] - t0[2], ny = t1[1] - t0[1];
    parallelFor(nz * ny, [&](int i) {
	const int tz = t0[2] + i / ny, ty = t0[1] + i % ny;
	for (int tx = t0[0]; tx < t1[0]; ++tx)
	{
	    const int lin = a.indexToLinearTile(tx << 4, ty << 4, tz << 4);
	    fn(a.getLinearTile(lin), tx << 4, ty << 4, tz << 4);
	}
    });
}

inline Box clampBox(const Box &b, const UT_Vector3I &res, int grow)
{
    Box o;
    for (int a = 0; a < 3; ++a)
    {
	o.lo[a] = std::max<int64_t>(0, b.lo[a] - grow);
	o.hi[a] = std::min<int64_t>(res[a], b.hi[a] + grow);
    }
    return o;
}

// copies voxels of [box] into the dense image (other cells keep their value)
template <typename T>
inline void flatten(Dense<T> &out, const UT_VoxelArray<T> &a, const Box &box)
{
    forTilesInBox(a, box, [&](const UT_VoxelTile<T> *tile, int x0, int y0, int z0) {
	const int nx = tile->xres(), ny = tile->yres(), nz = tile->zres();
	const bool isConst = tile->isConstant();
	const T cv = isConst ? (*tile)(0, 0, 0) : T(0);
	for (int z = 0; z < nz; ++z)
	{
	    if (z0 + z < box.lo[2] || z0 + z >= box.hi[2]) continue;
	    for (int y = 0; y < ny; ++y)
	    {
		if (y0 + y < box.lo[1] || y0 + y >= box.hi[1]) continue;
		const int xa = int(std::max<int64_t>(0, box.lo[0] - x0)), xb = int(std::min<int64_t>(nx, box.hi[0] - x0));
		T *row = &out.at(x0, y0 + y, z0 + z);
		if (isConst) for (int x = xa; x < xb; ++x) row[x] = cv;
		else for (int x = xa; x < xb; ++x) row[x] = (*tile)(x, y, z);
	    }
	}
    });
}

// writes dense values of [box] back into the voxel array where keep(x,y,z) holds
template <typename T, typename Keep>
inline void unflatten(UT_VoxelArray<T> &a, Dense<T> &in, const Box &box, const Keep &keep)
{
    forTilesInBox(a, box, [&](const UT_VoxelTile<T> *tile, int x0, int y0, int z0) {
	const int nx = tile->xres(), ny = tile->yres(), nz = tile->zres();
	for (int z = 0; z < nz; ++z)
	{
	    if (z0 + z < box.lo[2] || z0 + z >= box.hi[2]) continue;
	    for (int y = 0; y < ny; ++y)
	    {
		if (y0 + y < box.lo[1] || y0 + y >= box.hi[1]) continue;
		const int xa = int(std::max<int64_t>(0, box.lo[0] - x0)), xb = int(std::min<int64_t>(nx, box.hi[0] - x0));
		for (int x = xa; x < xb; ++x)
		    if (keep(x0 + x, y0 + y, z0 + z)) a.setValue(x0 + x, y0 + y, z0 + z, in.at(x0 + x, y0 + y, z0 + z));
	    }
	}
    });
}

// [lo,hi) bounds of the voxels for which pred(value) holds; constant tiles are decided from one voxel
template <typename T, typename Pred>
inline Box boundsWhere(const UT_VoxelArray<T> &a, const Pred &pred)
{
    const UT_Vector3I res = a.getVoxelRes();
    Box all;
    for (int k = 0; k < 3; ++k) all.hi[k] = res[k];
    const int nThreads = 16;
    std::vector<std::array<int64_t, 6>> part(nThreads, std::array<int64_t, 6>{res[0], res[1], res[2], -1, -1, -1});
    const int nt = a.numTiles();
    parallelFor(nThreads, [&](int t) {
	auto &b = part[t];
	for (int i = t; i < nt; i += nThreads)
	{
	    const UT_VoxelTile<T> *tile = a.getLinearTile(i);
	    int tx, ty, tz;
	    a.linearTileToXYZ(i, tx, ty, tz);
	    const int nx = tile->xres(), ny = tile->yres(), nz = tile->zres();
	    auto grow = [&](int64_t x, int64_t y, int64_t z) {
		b[0] = std::min(b[0], x); b[1] = std::min(b[1], y); b[2] = std::min(b[2], z);
		b[3] = std::max(b[3], x); b[4] = std::max(b[4], y); b[5] = std::max(b[5], z);
	    };
	    if (tile->isConstant())
	    {
		if (pred((*tile)(0, 0, 0))) { grow(tx * 16, ty * 16, tz * 16); grow(tx * 16 + nx - 1, ty * 16 + ny - 1, tz * 16 + nz - 1); }
		continue;
	    }
	    for (int z = 0; z < nz; ++z)
		for (int y = 0; y < ny; ++y)
		    for (int x = 0; x < nx; ++x)
			if (pred((*tile)(x, y, z))) grow(tx * 16 + x, ty * 16 + y, tz * 16 + z);
	}
    });
    Box out;
    int64_t b[6] = {res[0], res[1], res[2], -1, -1, -1};
    for (auto &p : part)
	for (int k = 0; k < 3; ++k) { b[k] = std::min(b[k], p[k]); b[k + 3] = std::max(b[k + 3], p[k + 3]); }
    if (b[3] < 0) return out;
    for (int k = 0; k < 3; ++k) { out.lo[k] = b[k]; out.hi[k] = b[k + 3] + 1; }
    return out;
}
} // namespace B200

// ----------------------------------------------------------------------------------------------------
// HDK::GeometricMultigridOperators (HDK_GeometricMultigridOperators.h:8-174)
// ----------------------------------------------------------------------------------------------------
namespace GeometricMultigridOperators
{
enum CellLabels { INTERIOR_CELL, EXTERIOR_CELL, DIRICHLET_CELL, BOUNDARY_CELL }; // Ops.h:11

namespace detail
{
template <typename Real>
struct OnlyDouble
{
    static_assert(std::is_same<Real, double>::value, "the B200 path is fp64 only, like the reference's solver (MG.h:14-15)");
};
inline bool isActive(int l) { return l == INTERIOR_CELL || l == BOUNDARY_CELL; }
} // namespace detail

// A label grid (+ optional level-0 face weights) resident on the device with its boundary band and
// coefficient records: what every stateless operator below needs.  levels > 1 also builds the coarse label
// grids (for downsample / upsampleAndAdd).
class DeviceDomain
{
public:
    DeviceDomain(const UT_VoxelArray<int> &cellLabels, const std::array<UT_VoxelArray<double>, 3> *boundaryWeights, int levels = 1,
		 bool operatorsOnly = true, bool useGaussSeidel = false, bool doPrintStats = false)
    {
	res = cellLabels.getVoxelRes();
	box = B200::boundsWhere(cellLabels, [](int l) { return l != EXTERIOR_CELL; });
	if (box.empty()) throw B200::Error(GMG_ERR_NO_ACTIVE, "DeviceDomain: the label grid has no non-EXTERIOR cell");
	const B200::Box io = B200::clampBox(box, res, 4);
	labels.reset(res);
	B200::flatten(labels, cellLabels, io);
	// outside the copied region the library never reads (box hint below), so the zero pages are never seen
	B200::Dense<double> w[3];
	if (boundaryWeights)
	    for (int a = 0; a < 3; ++a)
	    {
		UT_Vector3I fr = res;
		fr[a] += 1;
		if (!((*boundaryWeights)[a].getVoxelRes() == fr)) throw B200::Error(GMG_ERR_INVALID, "boundary weight grid shape (MG.cpp:167-177)");
		w[a].reset(fr);
		B200::Box fio = io;
		fio.hi[a] = std::min<int64_t>(fr[a], fio.hi[a] + 1);
		B200::flatten(w[a], (*boundaryWeights)[a], fio);
	    }
	gmg_solver_options opt;
	gmg_solver_default_options(&opt);
	opt.use_gauss_seidel = useGaussSeidel;
	opt.print_stats = doPrintStats;
	opt.operators_only = operatorsOnly;
	for (int a = 0; a < 3; ++a) { opt.box_lo[a] = box.lo[a]; opt.box_hi[a] = box.hi[a]; }
	const int64_t r[3] = {res[0], res[1], res[2]};
	B200::check(gmg_solver_create(B200::context(), labels.data, r, w[0].data, w[1].data, w[2].data, levels, &opt, &solver), "gmg_solver_create");
    }
    ~DeviceDomain() { gmg_solver_destroy(solver); }
    DeviceDomain(const DeviceDomain &) = delete;
    DeviceDomain &operator=(const DeviceDomain &) = delete;

    struct Grid
    {
	gmg_grid *g = nullptr;
	Grid() {}
	Grid(Grid &&o) noexcept : g(o.g) { o.g = nullptr; }
	Grid(const Grid &) = delete;
	~Grid() { gmg_grid_destroy(g); }
    };

    UT_Vector3I levelRes(int level) const
    {
	int64_t r[3];
	B200::check(gmg_solver_level_res(solver, level, r), "gmg_solver_level_res");
	return UT_Vector3I(r[0], r[1], r[2]);
    }
    B200::Box levelBox(int level) const
    {
	B200::Box b = box;
	for (int l = 0; l < level; ++l)
	    for (int a = 0; a < 3; ++a) { b.lo[a] = b.lo[a] >> 1; b.hi[a] = (b.hi[a] + 1) >> 1; }
	return b;
    }
    Grid upload(const UT_VoxelArray<double> &v, int level = 0) const
    {
	const UT_Vector3I r = levelRes(level);
	if (!(v.getVoxelRes() == r)) throw B200::Error(GMG_ERR_INVALID, "vector grid resolution does not match the label grid");
	B200::Dense<double> d(r);
	B200::flatten(d, v, B200::clampBox(levelBox(level), r, 4));
	Grid g;
	B200::check(gmg_grid_create(solver, level, &g.g), "gmg_grid_create");
	B200::check(gmg_grid_upload(g.g, d.data), "gmg_grid_upload");
	return g;
    }
    Grid zeros(int level = 0) const
    {
	Grid g;
	B200::check(gmg_grid_create(solver, level, &g.g), "gmg_grid_create");
	return g;
    }
    // active cells of the level get the device values; every other voxel of `v` is left as it is (the reference's
    // operators only ever write active cells)
    void download(UT_VoxelArray<double> &v, const Grid &g, int level = 0) const
    {
	const UT_Vector3I r = levelRes(level);
	B200::Dense<double> d(r);
	B200::check(gmg_grid_download(g.g, d.data), "gmg_grid_download");
	if (level == 0)
	{
	    B200::Dense<int> &lab = const_cast<B200::Dense<int> &>(labels);
	    B200::unflatten(v, d, box, [&](int64_t x, int64_t y, int64_t z) { return detail::isActive(lab.at(x, y, z)); });
	}
	else
	{
	    B200::Dense<int> lab(r);
	    B200::check(gmg_solver_get_labels(solver, level, lab.data), "gmg_solver_get_labels");
	    B200::unflatten(v, d, levelBox(level), [&](int64_t x, int64_t y, int64_t z) { return detail::isActive(lab.at(x, y, z)); });
	}
    }

    gmg_solver *solver = nullptr;
    UT_Vector3I res;
    B200::Box box;
    B200::Dense<int> labels;
};

// ---- smoothers, operator, residual (Ops.h:262-732) ---------------------------------------------------------
template <typename SolveReal, typename StoreReal>
void jacobiPoissonSmoother(UT_VoxelArray<StoreReal> &solution, const UT_VoxelArray<StoreReal> &rhs, const UT_VoxelArray<int> &cellLabels,
			   const std::array<UT_VoxelArray<StoreReal>, 3> *boundaryWeights = nullptr)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, boundaryWeights);
    auto x = d.upload(solution), b = d.upload(rhs);
    B200::check(gmg_jacobi(d.solver, x.g, b.g), "gmg_jacobi");
    d.download(solution, x);
}

// HDK_GeometricMultigridOperators.h:369-520: one half-pass over the odd or even 16^3 tiles, forwards or backwards
template <typename SolveReal, typename StoreReal>
void tiledGaussSeidelPoissonSmoother(UT_VoxelArray<StoreReal> &solution, const UT_VoxelArray<StoreReal> &rhs, const UT_VoxelArray<int> &cellLabels,
				     const bool doSmoothOddTiles, const bool doSmoothForward,
				     const std::array<UT_VoxelArray<StoreReal>, 3> *boundaryWeights = nullptr)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, boundaryWeights, 1, true, true);
    auto x = d.upload(solution), b = d.upload(rhs);
    B200::check(gmg_gauss_seidel(d.solver, x.g, b.g, doSmoothOddTiles ? 1 : 0, doSmoothForward ? 1 : 0), "gmg_gauss_seidel");
    d.download(solution, x);
}

// The band must be the one buildBoundaryCells(cellLabels, 3) returns (what every caller in the reference passes,
// MG.cpp:279-281); any other list is rejected rather than silently smoothed differently.
template <typename SolveReal, typename StoreReal>
void boundaryJacobiPoissonSmoother(UT_VoxelArray<StoreReal> &solution, const UT_VoxelArray<StoreReal> &rhs, const UT_VoxelArray<int> &cellLabels,
				   const UT_Array<UT_Vector3I> &boundaryCells, const std::array<UT_VoxelArray<StoreReal>, 3> *boundaryWeights = nullptr)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, boundaryWeights);
    int64_t n = 0;
    B200::check(gmg_solver_get_boundary_cells(d.solver, 0, nullptr, &n), "gmg_solver_get_boundary_cells");
    if (n != int64_t(boundaryCells.size())) throw B200::Error(GMG_ERR_INVALID, "boundaryJacobiPoissonSmoother: the cell list is not buildBoundaryCells(cellLabels, 3)");
    std::vector<int64_t> own(size_t(3 * std::max<int64_t>(n, 1)));
    B200::check(gmg_solver_get_boundary_cells(d.solver, 0, own.data(), &n), "gmg_solver_get_boundary_cells");
    for (int64_t i = 0; i < n; ++i)
	for (int a = 0; a < 3; ++a)
	    if (own[3 * i + a] != boundaryCells[i][a])
		throw B200::Error(GMG_ERR_INVALID, "boundaryJacobiPoissonSmoother: the cell list is not buildBoundaryCells(cellLabels, 3)");
    auto x = d.upload(solution), b = d.upload(rhs);
    B200::check(gmg_boundary_jacobi(d.solver, x.g, b.g, 1), "gmg_boundary_jacobi");
    d.download(solution, x);
}

template <typename SolveReal, typename StoreReal>
void applyPoissonMatrix(UT_VoxelArray<StoreReal> &destination, const UT_VoxelArray<StoreReal> &source, const UT_VoxelArray<int> &cellLabels,
			const std::array<UT_VoxelArray<StoreReal>, 3> *boundaryWeights = nullptr)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, boundaryWeights);
    auto src = d.upload(source), dst = d.zeros();
    B200::check(gmg_apply(d.solver, dst.g, src.g), "gmg_apply");
    d.download(destination, dst);
}

template <typename SolveReal, typename StoreReal>
void computePoissonResidual(UT_VoxelArray<StoreReal> &residual, const UT_VoxelArray<StoreReal> &solution, const UT_VoxelArray<StoreReal> &rhs,
			    const UT_VoxelArray<int> &cellLabels, const std::array<UT_VoxelArray<StoreReal>, 3> *boundaryWeights = nullptr)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, boundaryWeights);
    auto x = d.upload(solution), b = d.upload(rhs), r = d.zeros();
    B200::check(gmg_residual(d.solver, r.g, x.g, b.g), "gmg_residual");
    residual.constant(0); // Ops.h:726
    d.download(residual, r);
}

// ---- transfer operators (Ops.h:734-972).  The coarse labels must be buildCoarseCellLabels(fine labels). ----------
namespace detail
{
inline void checkCoarseLabels(const DeviceDomain &d, const UT_VoxelArray<int> &coarseLabels)
{
    const UT_Vector3I r = d.levelRes(1);
    if (!(coarseLabels.getVoxelRes() == r)) throw B200::Error(GMG_ERR_INVALID, "coarse label grid must be half the fine resolution");
    B200::Dense<int> own(r);
    B200::check(gmg_solver_get_labels(d.solver, 1, own.data), "gmg_solver_get_labels");
    const B200::Box b = B200::clampBox(d.levelBox(1), r, 1);
    for (int64_t z = b.lo[2]; z < b.hi[2]; ++z)
	for (int64_t y = b.lo[1]; y < b.hi[1]; ++y)
	    for (int64_t x = b.lo[0]; x < b.hi[0]; ++x)
		if (own.at(x, y, z) != coarseLabels(int(x), int(y), int(z)))
		    throw B200::Error(GMG_ERR_INVALID, "transfer operator: the coarse labels are not buildCoarseCellLabels(fine labels)");
}
} // namespace detail

template <typename SolveReal, typename StoreReal>
void downsample(UT_VoxelArray<StoreReal> &destination, const UT_VoxelArray<StoreReal> &source, const UT_VoxelArray<int> &destinationCellLabels,
		const UT_VoxelArray<int> &sourceCellLabels)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(sourceCellLabels, nullptr, 2);
    detail::checkCoarseLabels(d, destinationCellLabels);
    auto fine = d.upload(source), coarse = d.zeros(1);
    B200::check(gmg_restrict(d.solver, coarse.g, fine.g), "gmg_restrict");
    destination.constant(0); // Ops.h:756
    d.download(destination, coarse, 1);
}

template <typename SolveReal, typename StoreReal>
void upsampleAndAdd(UT_VoxelArray<StoreReal> &destination, const UT_VoxelArray<StoreReal> &source, const UT_VoxelArray<int> &destinationCellLabels,
		    const UT_VoxelArray<int> &sourceCellLabels)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(destinationCellLabels, nullptr, 2);
    detail::checkCoarseLabels(d, sourceCellLabels);
    auto fine = d.upload(destination), coarse = d.upload(source, 1);
    B200::check(gmg_prolong_add(d.solver, fine.g, coarse.g), "gmg_prolong_add");
    d.download(destination, fine);
}

// ---- BLAS-1 (Ops.h:974-1326) ------------------------------------------------------------------------------------
template <typename SolveReal, typename StoreReal>
void addToVector(UT_VoxelArray<StoreReal> &destination, const UT_VoxelArray<StoreReal> &source, const SolveReal scale, const UT_VoxelArray<int> &cellLabels)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, nullptr);
    auto y = d.upload(destination), a = d.upload(source);
    B200::check(gmg_axpy(d.solver, y.g, a.g, double(scale)), "gmg_axpy");
    d.download(destination, y);
}

template <typename SolveReal, typename StoreReal>
void addVectors(UT_VoxelArray<StoreReal> &destination, const UT_VoxelArray<StoreReal> &source, const UT_VoxelArray<StoreReal> &scaledSource,
		const SolveReal scale, const UT_VoxelArray<int> &cellLabels)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, nullptr);
    auto a = d.upload(source), v = d.upload(scaledSource), y = d.zeros();
    B200::check(gmg_add_scaled(d.solver, y.g, a.g, v.g, double(scale)), "gmg_add_scaled");
    d.download(destination, y);
}

template <typename SolveReal, typename StoreReal>
void scaleVector(UT_VoxelArray<StoreReal> &vector, const SolveReal scale, const UT_VoxelArray<int> &cellLabels)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, nullptr);
    auto y = d.upload(vector);
    B200::check(gmg_scale(d.solver, y.g, double(scale)), "gmg_scale");
    d.download(vector, y);
}

template <typename SolveReal, typename StoreReal>
SolveReal dotProduct(const UT_VoxelArray<StoreReal> &vectorA, const UT_VoxelArray<StoreReal> &vectorB, const UT_VoxelArray<int> &cellLabels)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, nullptr);
    auto a = d.upload(vectorA), b = d.upload(vectorB);
    double out = 0;
    B200::check(gmg_dot(d.solver, a.g, b.g, &out), "gmg_dot");
    return SolveReal(out);
}

template <typename SolveReal, typename StoreReal>
SolveReal squaredL2Norm(const UT_VoxelArray<StoreReal> &vector, const UT_VoxelArray<int> &cellLabels)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, nullptr);
    auto a = d.upload(vector);
    double out = 0;
    B200::check(gmg_norm2(d.solver, a.g, &out), "gmg_norm2");
    return SolveReal(out);
}

template <typename SolveReal, typename StoreReal>
SolveReal l2Norm(const UT_VoxelArray<StoreReal> &vector, const UT_VoxelArray<int> &cellLabels)
{
    return SolveReal(std::sqrt(squaredL2Norm<SolveReal>(vector, cellLabels))); // Ops.h:1197-1203
}

template <typename StoreReal>
StoreReal infNorm(const UT_VoxelArray<StoreReal> &vector, const UT_VoxelArray<int> &cellLabels)
{
    detail::OnlyDouble<StoreReal>();
    DeviceDomain d(cellLabels, nullptr);
    auto a = d.upload(vector);
    double out = 0;
    B200::check(gmg_inf_norm(d.solver, a.g, &out), "gmg_inf_norm"); // max(v, 0) like Ops.h:1303-1312
    return StoreReal(out);
}

// ---- domain builders (Ops.h:1328-1644, Ops.cpp:23-469) ------------------------------------------------------------
template <typename IsExteriorCellFunctor, typename IsInteriorCellFunctor, typename IsDirichletCellFunctor>
std::pair<UT_Vector3I, int> buildExpandedCellLabels(UT_VoxelArray<int> &expandedCellLabels, const UT_VoxelArray<int> &baseCellLabels,
						    const IsExteriorCellFunctor &isExteriorCell, const IsInteriorCellFunctor &isInteriorCell,
						    const IsDirichletCellFunctor &isDirichletCell)
{
    const UT_Vector3I bres = baseCellLabels.getVoxelRes();
    const int64_t br[3] = {bres[0], bres[1], bres[2]};
    int64_t er[3], off[3];
    int levels = 0;
    B200::check(gmg_expand_dims(br, er, off, &levels), "gmg_expand_dims");
    // the caller's label vocabulary -> the solver's three input labels, through the caller's own functors (Ops.h:1404-1450)
    B200::Dense<int> base(bres);
    B200::Box all;
    for (int a = 0; a < 3; ++a) all.hi[a] = bres[a];
    B200::flatten(base, baseCellLabels, all);
    const int64_t n = base.count();
    B200::parallelFor(16, [&](int t) {
	for (int64_t i = n * t / 16; i < n * (t + 1) / 16; ++i)
	{
	    const int l = base.data[i];
	    base.data[i] = isExteriorCell(l) ? EXTERIOR_CELL : (isInteriorCell(l) ? INTERIOR_CELL : (isDirichletCell(l) ? DIRICHLET_CELL : EXTERIOR_CELL));
	}
    });
    B200::Dense<int> out(UT_Vector3I(er[0], er[1], er[2]));
    B200::check(gmg_expand_labels(B200::context(), base.data, br, out.data, er, off), "gmg_expand_labels");
    expandedCellLabels.size(int(er[0]), int(er[1]), int(er[2]));
    expandedCellLabels.constant(EXTERIOR_CELL);
    B200::Box box;
    for (int a = 0; a < 3; ++a) { box.lo[a] = off[a]; box.hi[a] = off[a] + br[a]; }
    B200::unflatten(expandedCellLabels, out, box, [&](int64_t x, int64_t y, int64_t z) { return out.at(x, y, z) != EXTERIOR_CELL; });
    return std::pair<UT_Vector3I, int>(UT_Vector3I(off[0], off[1], off[2]), levels);
}

template <typename StoreReal>
void buildExpandedBoundaryWeights(UT_VoxelArray<StoreReal> &expandedBoundaryWeights, const UT_VoxelArray<StoreReal> &baseBoundaryWeights,
				  const UT_VoxelArray<int> &expandedCellLabels, const UT_Vector3I &exteriorOffset, const int axis)
{
    detail::OnlyDouble<StoreReal>();
    const UT_Vector3I fres = baseBoundaryWeights.getVoxelRes(), eres = expandedCellLabels.getVoxelRes();
    int64_t br[3] = {fres[0], fres[1], fres[2]};
    br[axis] -= 1;
    const int64_t er[3] = {eres[0], eres[1], eres[2]}, off[3] = {exteriorOffset[0], exteriorOffset[1], exteriorOffset[2]};
    B200::Dense<double> base(fres);
    B200::Box all;
    for (int a = 0; a < 3; ++a) all.hi[a] = fres[a];
    B200::flatten(base, baseBoundaryWeights, all);
    UT_Vector3I efr = eres;
    efr[axis] += 1;
    B200::Dense<double> out(efr);
    B200::check(gmg_expand_weights(B200::context(), base.data, br, out.data, er, off, axis), "gmg_expand_weights");
    expandedBoundaryWeights.size(int(efr[0]), int(efr[1]), int(efr[2]));
    expandedBoundaryWeights.constant(0);
    B200::Box box;
    for (int a = 0; a < 3; ++a) { box.lo[a] = off[a]; box.hi[a] = off[a] + fres[a]; }
    B200::unflatten(expandedBoundaryWeights, out, box, [&](int64_t x, int64_t y, int64_t z) { return out.at(x, y, z) != 0; });
}

template <typename StoreReal>
void setBoundaryCellLabels(UT_VoxelArray<int> &cellLabels, const std::array<UT_VoxelArray<StoreReal>, 3> &boundaryWeights)
{
    detail::OnlyDouble<StoreReal>();
    const UT_Vector3I res = cellLabels.getVoxelRes();
    const B200::Box box = B200::boundsWhere(cellLabels, [](int l) { return l != EXTERIOR_CELL; });
    if (box.empty()) return;
    const B200::Box io = B200::clampBox(box, res, 4);
    B200::Dense<int> lab(res);
    B200::flatten(lab, cellLabels, io);
    B200::Dense<double> w[3];
    for (int a = 0; a < 3; ++a)
    {
	UT_Vector3I fr = res;
	fr[a] += 1;
	w[a].reset(fr);
	B200::Box fio = io;
	fio.hi[a] = std::min<int64_t>(fr[a], fio.hi[a] + 1);
	B200::flatten(w[a], boundaryWeights[a], fio);
    }
    const int64_t r[3] = {res[0], res[1], res[2]};
    B200::check(gmg_set_boundary_labels(B200::context(), lab.data, r, w[0].data, w[1].data, w[2].data, box.lo, box.hi), "gmg_set_boundary_labels");
    B200::unflatten(cellLabels, lab, box, [&](int64_t x, int64_t y, int64_t z) { return lab.at(x, y, z) == BOUNDARY_CELL; });
}

inline UT_VoxelArray<int> buildCoarseCellLabels(const UT_VoxelArray<int> &sourceCellLabels)
{
    const UT_Vector3I res = sourceCellLabels.getVoxelRes();
    const B200::Box box = B200::boundsWhere(sourceCellLabels, [](int l) { return l != EXTERIOR_CELL; });
    UT_VoxelArray<int> coarse;
    coarse.size(int(res[0] / 2), int(res[1] / 2), int(res[2] / 2));
    coarse.constant(EXTERIOR_CELL);
    if (box.empty()) return coarse;
    B200::Dense<int> fine(res);
    std::fill(fine.data, fine.data + fine.count(), int(EXTERIOR_CELL)); // gmg_coarsen_labels scans the whole grid
    B200::flatten(fine, sourceCellLabels, box);
    B200::Dense<int> out(coarse.getVoxelRes());
    const int64_t r[3] = {res[0], res[1], res[2]};
    B200::check(gmg_coarsen_labels(B200::context(), fine.data, r, out.data), "gmg_coarsen_labels");
    B200::Box cb;
    for (int a = 0; a < 3; ++a) { cb.lo[a] = box.lo[a] >> 1; cb.hi[a] = (box.hi[a] + 1) >> 1; }
    B200::unflatten(coarse, out, cb, [&](int64_t x, int64_t y, int64_t z) { return out.at(x, y, z) != EXTERIOR_CELL; });
    return coarse;
}

inline UT_Array<UT_Vector3I> buildBoundaryCells(const UT_VoxelArray<int> &sourceCellLabels, const int boundaryWidth)
{
    const UT_Vector3I res = sourceCellLabels.getVoxelRes();
    const B200::Box box = B200::boundsWhere(sourceCellLabels, [](int l) { return l != EXTERIOR_CELL; });
    UT_Array<UT_Vector3I> cells;
    if (box.empty()) return cells;
    B200::Dense<int> lab(res);
    std::fill(lab.data, lab.data + lab.count(), int(EXTERIOR_CELL));
    B200::flatten(lab, sourceCellLabels, box);
    const int64_t r[3] = {res[0], res[1], res[2]};
    int64_t n = 0;
    B200::check(gmg_boundary_cells(B200::context(), lab.data, r, boundaryWidth, nullptr, &n), "gmg_boundary_cells");
    std::vector<int64_t> xyz(size_t(3 * std::max<int64_t>(n, 1)));
    B200::check(gmg_boundary_cells(B200::context(), lab.data, r, boundaryWidth, xyz.data(), &n), "gmg_boundary_cells");
    cells.setSize(n);
    for (int64_t i = 0; i < n; ++i) cells[i] = UT_Vector3I(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    return cells;
}

// Tile-decompression helpers (Ops.h:1646-1769): artefacts of UT_VoxelArray's constant-tile compression with no
// numeric effect; the device stores dense cropped boxes, so they are no-ops kept for source compatibility.
template <typename GridType>
void uncompressTiles(UT_VoxelArray<GridType> &, const UT_Array<bool> &) {}
template <typename GridType>
void uncompressBoundaryTiles(UT_VoxelArray<GridType> &, const UT_Array<UT_Vector3I> &) {}
template <typename GridType>
void uncompressActiveGrid(UT_VoxelArray<GridType> &, const UT_VoxelArray<int> &) {}
} // namespace GeometricMultigridOperators

// ----------------------------------------------------------------------------------------------------
// HDK::GeometricMultigridPoissonSolver (HDK_GeometricMultigridPoissonSolver.h:10-53)
// ----------------------------------------------------------------------------------------------------
class GeometricMultigridPoissonSolver
{
    using StoreReal = double;
    using SolveReal = double;

public:
    GeometricMultigridPoissonSolver(const UT_VoxelArray<int> &initialCellLabels, const std::array<UT_VoxelArray<StoreReal>, 3> &boundaryWeights,
				    const int mgLevels, const bool useGaussSeidel, const bool doPrintStats = false)
	: myDomain(new GeometricMultigridOperators::DeviceDomain(initialCellLabels, &boundaryWeights, mgLevels, false, useGaussSeidel, doPrintStats))
    {
    }

    // MG.cpp:420-881
    void applyVCycle(UT_VoxelArray<StoreReal> &solutionVector, const UT_VoxelArray<StoreReal> &rhsVector, const bool useInitialGuess = false)
    {
	auto b = myDomain->upload(rhsVector);
	auto x = useInitialGuess ? myDomain->upload(solutionVector) : myDomain->zeros();
	B200::check(gmg_vcycle_device(myDomain->solver, x.g, b.g, useInitialGuess), "gmg_vcycle_device");
	myDomain->download(solutionVector, x);
    }

    int getMGLevels()
    {
	int n = 0;
	B200::check(gmg_solver_levels(myDomain->solver, &n), "gmg_solver_levels");
	return n;
    }

    // ---- B200 additions ---------------------------------------------------------------------------------------
    // Whole PCG on the device (CG.h:11-207 wired as GFS.cpp:430-483: A = applyPoissonMatrix with the fine weights,
    // M^-1 = applyVCycle or identity).  Returns the iteration index CG.h:198 prints (-1 on the two early-outs).
    int solve(UT_VoxelArray<StoreReal> &solutionGrid, const UT_VoxelArray<StoreReal> &rhsGrid, StoreReal tolerance, int maxIterations,
	      bool useMGPreconditioner = true, std::vector<double> *relativeResidualHistory = nullptr)
    {
	// a solution grid that is still the constant zero of solutionGrid.constant(0) (GFS.cpp:392-398: no warm start) says so in O(tiles):
	// it is neither flattened nor uploaded
	auto b = myDomain->upload(rhsGrid), x = isConstantZero(solutionGrid) ? myDomain->zeros() : myDomain->upload(solutionGrid);
	std::vector<double> hist(size_t(maxIterations) + 2);
	int iterations = -1, count = 0;
	B200::check(gmg_pcg_device(myDomain->solver, x.g, b.g, tolerance, maxIterations, useMGPreconditioner, &iterations, hist.data(), int(hist.size()), &count),
		    "gmg_pcg_device");
	myDomain->download(solutionGrid, x);
	hist.resize(size_t(count));
	if (B200::verbose())
	{
	    for (double h : hist) std::cout << "    Relative error: " << h << std::endl;
	    std::cout << "Iterations: " << iterations << std::endl;
	}
	if (relativeResidualHistory) *relativeResidualHistory = hist;
	return iterations;
    }
    GeometricMultigridOperators::DeviceDomain &domain() { return *myDomain; }

    // every tile constant-compressed with the value 0 (UT_VoxelTile::isConstant / operator()): what constant(0) leaves behind
    static bool isConstantZero(const UT_VoxelArray<StoreReal> &v)
    {
	for (int i = 0, n = v.numTiles(); i < n; ++i)
	{
	    const auto *tile = v.getLinearTile(i);
	    if (!tile->isConstant() || (*tile)(0, 0, 0) != StoreReal(0)) return false;
	}
	return true;
    }

private:
    std::unique_ptr<GeometricMultigridOperators::DeviceDomain> myDomain;
};

// ----------------------------------------------------------------------------------------------------
// HDK::solveGeometricConjugateGradient (HDK_GeometricCGPoissonSolver.h:11-207)
// ----------------------------------------------------------------------------------------------------
// Fast path: everything on the device.
inline int solveGeometricConjugateGradient(GeometricMultigridPoissonSolver &solver, UT_VoxelArray<double> &solutionGrid, const UT_VoxelArray<double> &rhsGrid,
					   const double tolerance, const int maxIterations, bool useMGPreconditioner = true,
					   std::vector<double> *relativeResidualHistory = nullptr)
{
    return solver.solve(solutionGrid, rhsGrid, tolerance, maxIterations, useMGPreconditioner, relativeResidualHistory);
}

// Source-compatible functor form: the Krylov recurrence runs on the host over the caller's six functors (which may be
// the stateless operators above, the reference's CPU operators, or anything else).  For parity tests and odd callers;
// production callers use the overload above.
template <typename MatrixVectorMultiplyFunctor, typename PreconditionerFunctor, typename DotProductFunctor, typename SquaredL2NormFunctor,
	  typename AddToVectorFunctor, typename AddScaledVectorFunctor, typename StoreReal>
void solveGeometricConjugateGradient(UT_VoxelArray<StoreReal> &solutionGrid, const UT_VoxelArray<StoreReal> &rhsGrid,
				     const MatrixVectorMultiplyFunctor &matrixVectorMultiplyFunctor, const PreconditionerFunctor &preconditionerFunctor,
				     const DotProductFunctor &dotProductFunctor, const SquaredL2NormFunctor &squaredNormFunctor,
				     const AddToVectorFunctor &addToVectorFunctor, const AddScaledVectorFunctor &addScaledVectorFunctor,
				     const StoreReal tolerance, const int maxIterations)
{
    const UT_Vector3I res = solutionGrid.getVoxelRes();
    auto sized = [&](UT_VoxelArray<StoreReal> &g) { g.size(int(res[0]), int(res[1]), int(res[2])); g.constant(0); };
    const double bb = squaredNormFunctor(rhsGrid);
    if (bb == 0) { std::cout << "RHS is zero. Nothing to solve" << std::endl; return; }
    UT_VoxelArray<StoreReal> r, p, z, t;
    sized(r); sized(p); sized(z); sized(t);
    matrixVectorMultiplyFunctor(r, solutionGrid);
    addScaledVectorFunctor(r, rhsGrid, r, -1);
    double rr = squaredNormFunctor(r);
    const double threshold = double(tolerance) * double(tolerance) * bb;
    if (rr < threshold) { std::cout << "Residual already below error: " << std::sqrt(rr / bb) << std::endl; return; }
    preconditionerFunctor(p, r);
    double rho = dotProductFunctor(p, r);
    int iteration = 0;
    for (; iteration < maxIterations; ++iteration)
    {
	matrixVectorMultiplyFunctor(t, p);
	const double alpha = rho / dotProductFunctor(p, t);
	addToVectorFunctor(solutionGrid, p, alpha);
	addToVectorFunctor(r, t, -alpha);
	rr = squaredNormFunctor(r);
	std::cout << "    Relative error: " << std::sqrt(rr / bb) << std::endl;
	if (rr < threshold) break;
	preconditionerFunctor(z, r);
	const double rhoNew = dotProductFunctor(z, r);
	const double beta = rhoNew / rho;
	addScaledVectorFunctor(p, z, p, beta);
	rho = rhoNew;
    }
    std::cout << "Iterations: " << iteration << std::endl;
}
} // namespace GMG_HDK_NAMESPACE
