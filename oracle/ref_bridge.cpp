// ref_bridge.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" bridge over the reference's OWN hot-path sources (and HDK_Utilities.{h,cpp}), compiled
// unmodified from /root/reference/Source against oracle/shim (HDK + Eigen
// stand-ins; neither is installable offline).  Output: oracle/_ref/libgmg_ref.so.
// Used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  The product never links or loads it.
//
// Every entry point takes dense x-fastest arrays (idx = x + rx*(y + ry*z)) in the
// reference's own (expanded) coordinates, moves them into UT_VoxelArray tiles
// (collapsing constant tiles, as Houdini's grids are), calls the reference
// function named in the comment, and moves the result back.
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <vector>

#include "HDK_GeometricCGPoissonSolver.h"
#include "HDK_GeometricMultigridOperators.h"
#include "HDK_GeometricMultigridPoissonSolver.h"
#include "HDK_Utilities.h"
// the node's own source (HDK_GeometricFreeSurfacePressureSolver.cpp, compiled unmodified over shim/hdk_node_shim.h): its builders
// are private members, so this translation unit -- test infrastructure -- reads the class declaration with every member public.
// Everything the header pulls in has been included above already, so the two macros touch nothing but the class itself.
#include "hdk_node_shim.h"
#define private public
#define protected public
#include "HDK_GeometricFreeSurfacePressureSolver.h"
#undef protected
#undef private

namespace Ops = HDK::GeometricMultigridOperators;
using Real = double;
using WeightArray = std::array<UT_VoxelArray<Real>, 3>;

namespace
{
template <typename T>
void toVoxels(UT_VoxelArray<T> &dst, const T *src, const int64_t res[3], bool collapse = true)
{
    dst.size(int(res[0]), int(res[1]), int(res[2]));
    dst.constant(T(0));
    const int nt = dst.numTiles();
#pragma omp parallel for schedule(dynamic, 16)
    for (int t = 0; t < nt; ++t)
    {
	int tx, ty, tz;
	dst.linearTileToXYZ(t, tx, ty, tz);
	UT_VoxelTile<T> *tile = dst.getLinearTile(t);
	const int x0 = tx * 16, y0 = ty * 16, z0 = tz * 16;
	const int nx = tile->xres(), ny = tile->yres(), nz = tile->zres();
	// constant-tile test first to avoid allocating
	const T first = src[x0 + res[0] * (y0 + res[1] * int64_t(z0))];
	bool isConst = true;
	for (int z = 0; z < nz && isConst; ++z)
	    for (int y = 0; y < ny && isConst; ++y)
	    {
		const T *row = src + x0 + res[0] * ((y0 + y) + res[1] * int64_t(z0 + z));
		for (int x = 0; x < nx; ++x)
		    if (!(row[x] == first)) { isConst = false; break; }
	    }
	if (isConst && collapse) { tile->makeConstant(first); continue; }
	tile->uncompress();
	T *d = tile->rawData();
	for (int z = 0; z < nz; ++z)
	    for (int y = 0; y < ny; ++y)
		std::memcpy(d + (z * ny + y) * nx, src + x0 + res[0] * ((y0 + y) + res[1] * int64_t(z0 + z)), sizeof(T) * nx);
    }
}

template <typename T>
void fromVoxels(T *dst, const UT_VoxelArray<T> &src)
{
    const UT_Vector3I res = src.getVoxelRes();
    const int nt = src.numTiles();
#pragma omp parallel for schedule(dynamic, 16)
    for (int t = 0; t < nt; ++t)
    {
	int tx, ty, tz;
	src.linearTileToXYZ(t, tx, ty, tz);
	const UT_VoxelTile<T> *tile = src.getLinearTile(t);
	const int x0 = tx * 16, y0 = ty * 16, z0 = tz * 16;
	const int nx = tile->xres(), ny = tile->yres(), nz = tile->zres();
	for (int z = 0; z < nz; ++z)
	    for (int y = 0; y < ny; ++y)
	    {
		T *row = dst + x0 + res[0] * ((y0 + y) + res[1] * int64_t(z0 + z));
		if (tile->isConstant())
		    for (int x = 0; x < nx; ++x) row[x] = tile->constantValue();
		else
		    std::memcpy(row, tile->rawData() + (z * ny + y) * nx, sizeof(T) * nx);
	    }
    }
}

void faceRes(int64_t out[3], const int64_t res[3], int axis)
{
    out[0] = res[0]; out[1] = res[1]; out[2] = res[2];
    ++out[axis];
}

// weights may be null (coarse-level operators take no weights)
bool loadWeights(WeightArray &w, const double *w0, const double *w1, const double *w2, const int64_t res[3])
{
    if (!w0) return false;
    const double *ws[3] = {w0, w1, w2};
    for (int a = 0; a < 3; ++a)
    {
	int64_t fr[3];
	faceRes(fr, res, a);
	toVoxels(w[a], ws[a], fr);
    }
    return true;
}

struct CoutSilencer
{
    std::streambuf *old;
    std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};

double nowSeconds()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace

struct RefSolver
{
    UT_VoxelArray<int> labels;
    WeightArray weights;
    HDK::GeometricMultigridPoissonSolver *mg = nullptr;
    int64_t res[3];
    double setupSeconds = 0;
};

extern "C"
{
int ref_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int ref_jobs() { return UT_Thread::getNumProcessors(); }

void ref_free(void *p) { std::free(p); }

// Ops.h:1328-1456 buildExpandedCellLabels. *out is malloc'd (ref_free).
int ref_expand_labels(const int *base, const int64_t res[3], int **out, int64_t outRes[3], int64_t offset[3], int *mgLevels)
{
    UT_VoxelArray<int> baseLabels, expanded;
    toVoxels(baseLabels, base, res);
    auto isExt = [](const int v) { return v == Ops::EXTERIOR_CELL; };
    auto isInt = [](const int v) { return v == Ops::INTERIOR_CELL; };
    auto isDir = [](const int v) { return v == Ops::DIRICHLET_CELL; };
    std::pair<UT_Vector3I, int> r = Ops::buildExpandedCellLabels(expanded, baseLabels, isExt, isInt, isDir);
    for (int a = 0; a < 3; ++a) { offset[a] = r.first[a]; outRes[a] = expanded.getVoxelRes()[a]; }
    *mgLevels = r.second;
    *out = static_cast<int *>(std::malloc(sizeof(int) * size_t(outRes[0]) * outRes[1] * outRes[2]));
    fromVoxels(*out, expanded);
    return 0;
}

// Ops.h:1458-1572 buildExpandedBoundaryWeights. outW preallocated: expRes (+1 along axis).
int ref_expand_weights(const double *baseW, const int64_t baseRes[3], const int *expLabels, const int64_t expRes[3],
		       const int64_t offset[3], int axis, double *outW)
{
    int64_t bfr[3], efr[3];
    faceRes(bfr, baseRes, axis);
    faceRes(efr, expRes, axis);
    UT_VoxelArray<Real> bw, ew;
    UT_VoxelArray<int> labels;
    toVoxels(bw, baseW, bfr);
    toVoxels(labels, expLabels, expRes);
    ew.size(int(efr[0]), int(efr[1]), int(efr[2]));
    ew.constant(0);
    Ops::buildExpandedBoundaryWeights(ew, bw, labels, UT_Vector3I(offset[0], offset[1], offset[2]), axis);
    fromVoxels(outW, ew);
    return 0;
}

// Ops.h:1574-1644 setBoundaryCellLabels (labels in place)
int ref_set_boundary_labels(int *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2)
{
    UT_VoxelArray<int> l;
    // The reference writes cells of tiles it found non-constant-or-INTERIOR; a constant
    // INTERIOR tile would be written through setValue (asserted uncompressed in debug),
    // so hand it uncollapsed tiles exactly as buildExpandedCellLabels leaves them.
    toVoxels(l, labels, res, false);
    WeightArray w;
    loadWeights(w, w0, w1, w2, res);
    Ops::setBoundaryCellLabels(l, w);
    fromVoxels(labels, l);
    return 0;
}

// Ops.cpp:23-163 buildCoarseCellLabels. coarse preallocated (res/2).
int ref_coarsen_labels(const int *fine, const int64_t res[3], int *coarse)
{
    UT_VoxelArray<int> f;
    toVoxels(f, fine, res);
    UT_VoxelArray<int> c = Ops::buildCoarseCellLabels(f);
    fromVoxels(coarse, c);
    return 0;
}

// Ops.cpp:165-469 buildBoundaryCells. *xyz malloc'd int64[3*count] (ref_free).
int ref_boundary_cells(const int *labels, const int64_t res[3], int width, int64_t **xyz, int64_t *count)
{
    UT_VoxelArray<int> l;
    toVoxels(l, labels, res);
    UT_Array<UT_Vector3I> cells = Ops::buildBoundaryCells(l, width);
    *count = cells.size();
    *xyz = static_cast<int64_t *>(std::malloc(sizeof(int64_t) * 3 * size_t(cells.size() + 1)));
    for (exint i = 0; i < cells.size(); ++i)
	for (int a = 0; a < 3; ++a) (*xyz)[3 * i + a] = cells[i][a];
    return 0;
}

// Invariant checkers: Ops.h:1771-1870, Ops.cpp:602-632, Ops.cpp:471-600
int ref_unit_test_boundary_cells(const int *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2)
{
    UT_VoxelArray<int> l;
    toVoxels(l, labels, res);
    WeightArray w;
    const bool hasW = loadWeights(w, w0, w1, w2, res);
    return Ops::unitTestBoundaryCells<Real>(l, hasW ? &w : nullptr) ? 1 : 0;
}
int ref_unit_test_exterior_cells(const int *labels, const int64_t res[3])
{
    UT_VoxelArray<int> l;
    toVoxels(l, labels, res);
    return Ops::unitTestExteriorCells(l) ? 1 : 0;
}
int ref_unit_test_coarsening(const int *coarse, const int *fine, const int64_t fineRes[3])
{
    int64_t cres[3] = {fineRes[0] / 2, fineRes[1] / 2, fineRes[2] / 2};
    UT_VoxelArray<int> c, f;
    toVoxels(c, coarse, cres);
    toVoxels(f, fine, fineRes);
    return Ops::unitTestCoarsening(c, f) ? 1 : 0;
}

// ---- grid operators ---------------------------------------------------------
// Ops.h:262-367 jacobiPoissonSmoother (x in place)
int ref_jacobi(double *x, const double *b, const int *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2)
{
    UT_VoxelArray<Real> X, B;
    UT_VoxelArray<int> L;
    toVoxels(X, x, res); toVoxels(B, b, res); toVoxels(L, labels, res);
    WeightArray w;
    const bool hasW = loadWeights(w, w0, w1, w2, res);
    Ops::uncompressActiveGrid(X, L);
    Ops::jacobiPoissonSmoother<Real>(X, B, L, hasW ? &w : nullptr);
    fromVoxels(x, X);
    return 0;
}

// Ops.h:369-520 tiledGaussSeidelPoissonSmoother (x in place)
int ref_gauss_seidel(double *x, const double *b, const int *labels, const int64_t res[3], int oddTiles, int forward,
		     const double *w0, const double *w1, const double *w2)
{
    UT_VoxelArray<Real> X, B;
    UT_VoxelArray<int> L;
    toVoxels(X, x, res); toVoxels(B, b, res); toVoxels(L, labels, res);
    WeightArray w;
    const bool hasW = loadWeights(w, w0, w1, w2, res);
    Ops::uncompressActiveGrid(X, L);
    Ops::tiledGaussSeidelPoissonSmoother<Real>(X, B, L, oddTiles != 0, forward != 0, hasW ? &w : nullptr);
    fromVoxels(x, X);
    return 0;
}

// Ops.h:524-619 boundaryJacobiPoissonSmoother (x in place), `sweeps` applications
int ref_boundary_jacobi(double *x, const double *b, const int *labels, const int64_t res[3], const int64_t *cells, int64_t count,
			int sweeps, const double *w0, const double *w1, const double *w2)
{
    UT_VoxelArray<Real> X, B;
    UT_VoxelArray<int> L;
    toVoxels(X, x, res); toVoxels(B, b, res); toVoxels(L, labels, res);
    WeightArray w;
    const bool hasW = loadWeights(w, w0, w1, w2, res);
    UT_Array<UT_Vector3I> list;
    list.setSize(count);
    for (int64_t i = 0; i < count; ++i) list[i] = UT_Vector3I(cells[3 * i], cells[3 * i + 1], cells[3 * i + 2]);
    Ops::uncompressBoundaryTiles(X, list);
    for (int s = 0; s < sweeps; ++s)
	Ops::boundaryJacobiPoissonSmoother<Real>(X, B, L, list, hasW ? &w : nullptr);
    fromVoxels(x, X);
    return 0;
}

// Ops.h:621-714 applyPoissonMatrix (dst written on active cells only; dst comes in as given)
int ref_apply(double *dst, const double *src, const int *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2)
{
    UT_VoxelArray<Real> D, S;
    UT_VoxelArray<int> L;
    toVoxels(D, dst, res); toVoxels(S, src, res); toVoxels(L, labels, res);
    WeightArray w;
    const bool hasW = loadWeights(w, w0, w1, w2, res);
    Ops::applyPoissonMatrix<Real>(D, S, L, hasW ? &w : nullptr);
    fromVoxels(dst, D);
    return 0;
}

// Ops.h:716-732 computePoissonResidual
int ref_residual(double *r, const double *x, const double *b, const int *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2)
{
    UT_VoxelArray<Real> R, X, B;
    UT_VoxelArray<int> L;
    toVoxels(X, x, res); toVoxels(B, b, res); toVoxels(L, labels, res);
    R.size(int(res[0]), int(res[1]), int(res[2]));
    R.constant(0);
    WeightArray w;
    const bool hasW = loadWeights(w, w0, w1, w2, res);
    Ops::computePoissonResidual<Real>(R, X, B, L, hasW ? &w : nullptr);
    fromVoxels(r, R);
    return 0;
}

// Ops.h:734-835 downsample. coarseRes = fineRes/2.
int ref_downsample(double *coarse, const double *fine, const int *coarseLabels, const int *fineLabels, const int64_t fineRes[3])
{
    int64_t cres[3] = {fineRes[0] / 2, fineRes[1] / 2, fineRes[2] / 2};
    UT_VoxelArray<Real> C, F;
    UT_VoxelArray<int> CL, FL;
    toVoxels(F, fine, fineRes); toVoxels(CL, coarseLabels, cres); toVoxels(FL, fineLabels, fineRes);
    C.size(int(cres[0]), int(cres[1]), int(cres[2]));
    C.constant(0);
    Ops::downsample<Real>(C, F, CL, FL);
    fromVoxels(coarse, C);
    return 0;
}

// Ops.h:873-972 upsampleAndAdd (fine in place)
int ref_upsample_add(double *fine, const double *coarse, const int *fineLabels, const int *coarseLabels, const int64_t fineRes[3])
{
    int64_t cres[3] = {fineRes[0] / 2, fineRes[1] / 2, fineRes[2] / 2};
    UT_VoxelArray<Real> C, F;
    UT_VoxelArray<int> CL, FL;
    toVoxels(F, fine, fineRes); toVoxels(C, coarse, cres); toVoxels(CL, coarseLabels, cres); toVoxels(FL, fineLabels, fineRes);
    Ops::uncompressActiveGrid(F, FL);
    Ops::upsampleAndAdd<Real>(F, C, FL, CL);
    fromVoxels(fine, F);
    return 0;
}

// Ops.h:1020-1085 / :1205-1265 / :1267-1326
double ref_dot(const double *a, const double *b, const int *labels, const int64_t res[3])
{
    UT_VoxelArray<Real> A, B;
    UT_VoxelArray<int> L;
    toVoxels(A, a, res); toVoxels(B, b, res); toVoxels(L, labels, res);
    return Ops::dotProduct<Real>(A, B, L);
}
double ref_norm2(const double *a, const int *labels, const int64_t res[3])
{
    UT_VoxelArray<Real> A;
    UT_VoxelArray<int> L;
    toVoxels(A, a, res); toVoxels(L, labels, res);
    return Ops::squaredL2Norm<Real>(A, L);
}
double ref_inf_norm(const double *a, const int *labels, const int64_t res[3])
{
    UT_VoxelArray<Real> A;
    UT_VoxelArray<int> L;
    toVoxels(A, a, res); toVoxels(L, labels, res);
    return Ops::infNorm(A, L);
}
// Ops.h:1087-1137 addToVector: dst += s*src
int ref_axpy(double *dst, const double *src, double s, const int *labels, const int64_t res[3])
{
    UT_VoxelArray<Real> D, S;
    UT_VoxelArray<int> L;
    toVoxels(D, dst, res); toVoxels(S, src, res); toVoxels(L, labels, res);
    Ops::uncompressActiveGrid(D, L);
    Ops::addToVector<Real>(D, S, s, L);
    fromVoxels(dst, D);
    return 0;
}
// Ops.h:1139-1195 addVectors: dst = a + s*v
int ref_add_scaled(double *dst, const double *a, const double *v, double s, const int *labels, const int64_t res[3])
{
    UT_VoxelArray<Real> D, A, V;
    UT_VoxelArray<int> L;
    toVoxels(D, dst, res); toVoxels(A, a, res); toVoxels(V, v, res); toVoxels(L, labels, res);
    Ops::uncompressActiveGrid(D, L);
    Ops::addVectors<Real>(D, A, V, s, L);
    fromVoxels(dst, D);
    return 0;
}
// Ops.h:974-1018 scaleVector
int ref_scale(double *v, double s, const int *labels, const int64_t res[3])
{
    UT_VoxelArray<Real> V;
    UT_VoxelArray<int> L;
    toVoxels(V, v, res); toVoxels(L, labels, res);
    Ops::scaleVector<Real>(V, s, L);
    fromVoxels(v, V);
    return 0;
}

// ---- solver ------------------------------------------------------------------
// MG.cpp:135-418 constructor
RefSolver *ref_solver_create(const int *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2,
			     int mgLevels, int useGaussSeidel)
{
    CoutSilencer quiet;
    RefSolver *s = new RefSolver;
    for (int a = 0; a < 3; ++a) s->res[a] = res[a];
    toVoxels(s->labels, labels, res);
    loadWeights(s->weights, w0, w1, w2, res);
    const double t0 = nowSeconds();
    s->mg = new HDK::GeometricMultigridPoissonSolver(s->labels, s->weights, mgLevels, useGaussSeidel != 0, false);
    s->setupSeconds = nowSeconds() - t0;
    return s;
}
void ref_solver_destroy(RefSolver *s)
{
    if (!s) return;
    delete s->mg;
    delete s;
}
int ref_solver_levels(RefSolver *s) { return s->mg->getMGLevels(); }
double ref_solver_setup_seconds(RefSolver *s) { return s->setupSeconds; }

// MG.cpp:420-881 applyVCycle
double ref_solver_vcycle(RefSolver *s, double *x, const double *b, int useInitialGuess, int repeats)
{
    UT_VoxelArray<Real> X, B;
    toVoxels(X, x, s->res); toVoxels(B, b, s->res);
    Ops::uncompressActiveGrid(X, s->labels);
    const double t0 = nowSeconds();
    for (int r = 0; r < (repeats > 0 ? repeats : 1); ++r)
	s->mg->applyVCycle(X, B, useInitialGuess != 0);
    const double dt = nowSeconds() - t0;
    fromVoxels(x, X);
    return dt;
}

// CG.h:11-207 solveGeometricConjugateGradient wired exactly as Test.cpp:746-832 / GFS.cpp:430-483 do.
// precond: 1 = multigrid V-cycle (needs s->mg); 2 = the node's other mode, the diagonal preconditioner of GFS.cpp:485-603
// (that file needs live SIM fields and cannot be compiled here, so its two lambdas are restated below over the same
// containers: 1/6 on INTERIOR cells, 1/(sum of the six face weights) on BOUNDARY cells; destination = source * that).
// history[k] = sqrt(|r_k|^2/|b|^2) for every iteration the loop ran (the value CG.h:159 prints).
// Returns the iteration index CG.h:198 prints, or -1 on the two early-outs (CG.h:35-40, :60-64).
static int refPcgImpl(RefSolver *s, double *x, const double *b, double tol, int maxIt, int precond, double *history, int histCap, int *histCount,
		      double *solveSeconds)
{
    CoutSilencer quiet;
    UT_VoxelArray<Real> X, B;
    toVoxels(X, x, s->res); toVoxels(B, b, s->res);
    Ops::uncompressActiveGrid(X, s->labels);
    const UT_VoxelArray<int> &labels = s->labels;
    const WeightArray &weights = s->weights;
    HDK::GeometricMultigridPoissonSolver &mg = *s->mg;

    // GFS.cpp:487-560
    UT_VoxelArray<Real> diagonalPrecondGrid;
    if (precond == 2)
    {
	diagonalPrecondGrid.size(labels.getVoxelRes()[0], labels.getVoxelRes()[1], labels.getVoxelRes()[2]);
	diagonalPrecondGrid.constant(0);
	Ops::uncompressActiveGrid(diagonalPrecondGrid, labels);
	for (int64_t z = 0; z < s->res[2]; ++z)
	    for (int64_t y = 0; y < s->res[1]; ++y)
		for (int64_t xx = 0; xx < s->res[0]; ++xx)
		{
		    const int l = labels(xx, y, z);
		    UT_Vector3I cell(xx, y, z);
		    if (l == Ops::CellLabels::INTERIOR_CELL) diagonalPrecondGrid.setValue(cell, 1. / 6.);
		    else if (l == Ops::CellLabels::BOUNDARY_CELL)
		    {
			double diagonal = 0;
			for (int axis : {0, 1, 2})
			    for (int direction : {0, 1})
			    {
				UT_Vector3I face = SIM::FieldUtils::cellToFaceMap(cell, axis, direction);
				diagonal += weights[axis](face);
			    }
			diagonalPrecondGrid.setValue(cell, 1. / diagonal);
		    }
		}
    }

    std::vector<double> norms;
    auto A = [&](UT_VoxelArray<Real> &dst, const UT_VoxelArray<Real> &src) { Ops::applyPoissonMatrix<Real>(dst, src, labels, &weights); };
    auto M = [&](UT_VoxelArray<Real> &dst, const UT_VoxelArray<Real> &src) {
	if (precond == 1) { mg.applyVCycle(dst, src); return; }
	// GFS.cpp:562-603
	Ops::uncompressActiveGrid(dst, labels);
	for (int64_t z = 0; z < s->res[2]; ++z)
	    for (int64_t y = 0; y < s->res[1]; ++y)
		for (int64_t xx = 0; xx < s->res[0]; ++xx)
		{
		    const int l = labels(xx, y, z);
		    if (l == Ops::CellLabels::INTERIOR_CELL || l == Ops::CellLabels::BOUNDARY_CELL)
			dst.setValue(xx, y, z, double(src(xx, y, z)) * double(diagonalPrecondGrid(xx, y, z)));
		}
    };
    auto dot = [&](const UT_VoxelArray<Real> &a, const UT_VoxelArray<Real> &c) { return Ops::dotProduct<Real>(a, c, labels); };
    auto norm2 = [&](const UT_VoxelArray<Real> &a) { double v = Ops::squaredL2Norm<Real>(a, labels); norms.push_back(v); return v; };
    auto axpy = [&](UT_VoxelArray<Real> &dst, const UT_VoxelArray<Real> &src, const Real sc) { Ops::addToVector<Real>(dst, src, sc, labels); };
    auto addScaled = [&](UT_VoxelArray<Real> &dst, const UT_VoxelArray<Real> &a, const UT_VoxelArray<Real> &v, const Real sc) { Ops::addVectors<Real>(dst, a, v, sc, labels); };

    const double t0 = nowSeconds();
    HDK::solveGeometricConjugateGradient(X, B, A, M, dot, norm2, axpy, addScaled, Real(tol), maxIt);
    if (solveSeconds) *solveSeconds = nowSeconds() - t0;
    fromVoxels(x, X);

    // norm2 call sequence: |b|^2, |r0|^2, one per loop iteration, final recomputed.
    *histCount = 0;
    if (norms.size() < 3) return -1;
    const double rhsNorm2 = norms[0];
    const size_t loopCount = norms.size() - 3;
    for (size_t k = 0; k < loopCount && int(k) < histCap; ++k) history[(*histCount)++] = std::sqrt(norms[2 + k] / rhsNorm2);
    const double threshold = tol * tol * rhsNorm2;
    // CG.h prints the loop index at break; if the cap was hit it prints maxIt.
    if (loopCount > 0 && norms[1 + loopCount] < threshold) return int(loopCount) - 1;
    return int(loopCount);
}
int ref_pcg(RefSolver *s, double *x, const double *b, double tol, int maxIt, double *history, int histCap, int *histCount, double *solveSeconds)
{
    return refPcgImpl(s, x, b, tol, maxIt, 1, history, histCap, histCount, solveSeconds);
}
int ref_pcg_diag(RefSolver *s, double *x, const double *b, double tol, int maxIt, double *history, int histCap, int *histCount, double *solveSeconds)
{
    return refPcgImpl(s, x, b, tol, maxIt, 2, history, histCap, histCount, solveSeconds);
}
namespace
{
using Node = HDK_GeometricFreeSurfacePressureSolver;
void cellField(SIM_RawField &f, const float *src, const int64_t res[3])
{
    f.init(int(res[0]), int(res[1]), int(res[2]));
    if (src) toVoxels(*f.fieldNC(), src, res);
}
void faceField(SIM_RawField &f, const float *src, const int64_t res[3], int axis)
{
    f.init(SIM_FieldSample(SIM_SAMPLE_FACEX + axis), UT_Vector3(0, 0, 0), UT_Vector3(float(res[0]), float(res[1]), float(res[2])), int(res[0]), int(res[1]), int(res[2]));
    int64_t fr[3];
    faceRes(fr, res, axis);
    if (src) toVoxels(*f.fieldNC(), src, fr);
}
void vectorField(SIM_VectorField &v, const float *const src[3], const int64_t res[3])
{
    v.initFaces(int(res[0]), int(res[1]), int(res[2]));
    for (int a = 0; a < 3; ++a)
    {
	int64_t fr[3];
	faceRes(fr, res, a);
	if (src && src[a]) toVoxels(*v.getField(a)->fieldNC(), src[a], fr);
    }
}
void indexField(SIM_RawIndexField &f, const int32_t *src, const int64_t res[3])
{
    std::vector<exint> wide(size_t(res[0]) * res[1] * res[2]);
    for (size_t i = 0; i < wide.size(); ++i) wide[i] = src[i];
    toVoxels(*f.fieldNC(), wide.data(), res);
}
} // namespace

// ---- the steps in front of the solve whose reference source compiles here: HDK_Utilities.{h,cpp}, unmodified ----------------
// buildMaterialCellLabels, HDK_Utilities.cpp:87-148 (isCellLiquid :5-45).  Fields are cell-sampled on one lattice (oracle/shim:
// SIM_RawField), so solidSurface.getValue(indexToPos(cell)) is the solid field's own cell value.
void ref_build_material_labels(const float *liquidSurface, const float *solidSurface, const float *cut0, const float *cut1, const float *cut2, const int64_t res[3],
			       int32_t *material)
{
    SIM_RawField liquid, solid, cut[3];
    cellField(liquid, liquidSurface, res);
    cellField(solid, solidSurface, res);
    const float *cs[3] = {cut0, cut1, cut2};
    for (int a = 0; a < 3; ++a) faceField(cut[a], cs[a], res, a);
    SIM_RawIndexField labels;
    const std::array<const SIM_RawField *, 3> cutCellWeights = {&cut[0], &cut[1], &cut[2]};
    HDK::Utilities::buildMaterialCellLabels(labels, liquid, solid, cutCellWeights);
    std::vector<exint> wide(size_t(res[0]) * res[1] * res[2]);
    fromVoxels(wide.data(), *labels.field());
    for (size_t i = 0; i < wide.size(); ++i) material[i] = int32_t(wide[i]);
}
// buildValidFaces for one axis: the call sequence of HDK_GeometricFreeSurfacePressureSolver.cpp:717-744 (that file needs the node
// class and cannot be compiled) over the reference's own templates findOccupiedFaceTiles, uncompressTiles, classifyValidFaces
// (HDK_Utilities.h:78-189).
void ref_build_valid_faces(const int32_t *material, const float *cutCell, const int64_t res[3], int axis, float *validFaces)
{
    using MaterialLabels = HDK::Utilities::FreeSurfaceMaterialLabels;
    std::vector<exint> wide(size_t(res[0]) * res[1] * res[2]);
    for (size_t i = 0; i < wide.size(); ++i) wide[i] = material[i];
    SIM_RawIndexField labels;
    toVoxels(*labels.fieldNC(), wide.data(), res);
    SIM_RawField cut, valid;
    faceField(cut, cutCell, res, axis);
    faceField(valid, nullptr, res, axis);
    valid.makeConstant(HDK::Utilities::INVALID_FACE);
    UT_Array<bool> isTileOccupiedList;
    isTileOccupiedList.setSize(valid.field()->numTiles());
    isTileOccupiedList.constant(false);
    auto isLiquid = [](const exint label) { return label == MaterialLabels::LIQUID_CELL; };
    HDK::Utilities::findOccupiedFaceTiles(isTileOccupiedList, valid, labels, isLiquid, axis);
    HDK::Utilities::uncompressTiles(valid, isTileOccupiedList);
    HDK::Utilities::classifyValidFaces(valid, labels, cut, isLiquid, axis);
    fromVoxels(validFaces, *valid.field());
}
// ---- the node's own functions (HDK_GeometricFreeSurfacePressureSolver.cpp:746-1131 and solveGasSubclass :113-714), unmodified ------
// Fields live on one unit-spaced lattice: cell fields x-fastest [rz][ry][rx], the face field of axis a with one more entry along a.

// buildMGDomainLabels, :746-793 (the label grid starts as EXTERIOR, :309)
void ref_node_domain_labels(const int32_t *material, const int64_t res[3], int32_t *labels)
{
    Node node(nullptr);
    SIM_RawIndexField mat;
    indexField(mat, material, res);
    UT_VoxelArray<int> out;
    out.size(int(res[0]), int(res[1]), int(res[2]));
    out.constant(Ops::CellLabels::EXTERIOR_CELL);
    node.buildMGDomainLabels(out, mat);
    fromVoxels(labels, out);
}
// buildMGBoundaryWeights, :796-865, one axis (the weight grid starts as 0, :322)
void ref_node_boundary_weights(const float *cutCell, const float *liquidSurface, const float *validFaces, const int32_t *material, const int32_t *domainLabels,
			       const int64_t res[3], int axis, double *weights)
{
    Node node(nullptr);
    SIM_RawField cut, liquid, valid;
    faceField(cut, cutCell, res, axis);
    faceField(valid, validFaces, res, axis);
    cellField(liquid, liquidSurface, res);
    SIM_RawIndexField mat;
    indexField(mat, material, res);
    UT_VoxelArray<int> labels;
    toVoxels(labels, domainLabels, res);
    int64_t fr[3];
    faceRes(fr, res, axis);
    UT_VoxelArray<double> w;
    w.size(int(fr[0]), int(fr[1]), int(fr[2]));
    w.constant(0);
    node.buildMGBoundaryWeights(w, cut, liquid, valid, mat, labels, axis);
    fromVoxels(weights, w);
}
// buildRHS, :868-943: rhs is the expanded grid (zero on entry, :385); expLabels the expanded labels WITH their BOUNDARY cells set
void ref_node_rhs(const int32_t *material, const float *const velocity[3], const float *const cutCell[3], const float *const solidVelocity[3], const int64_t res[3],
		  const int32_t *expLabels, const int64_t expRes[3], const int64_t offset[3], double *rhs)
{
    Node node(nullptr);
    SIM_RawIndexField mat;
    indexField(mat, material, res);
    SIM_VectorField vel, solidVel;
    vectorField(vel, velocity, res);
    if (solidVelocity) vectorField(solidVel, solidVelocity, res);
    SIM_RawField cut[3];
    for (int a = 0; a < 3; ++a) faceField(cut[a], cutCell[a], res, a);
    const std::array<const SIM_RawField *, 3> cutCellWeights = {&cut[0], &cut[1], &cut[2]};
    UT_VoxelArray<int> labels;
    toVoxels(labels, expLabels, expRes);
    UT_VoxelArray<double> grid;
    grid.size(int(expRes[0]), int(expRes[1]), int(expRes[2]));
    grid.constant(0);
    node.buildRHS(grid, mat, vel, solidVelocity ? &solidVel : nullptr, cutCellWeights, labels, UT_Vector3I(offset[0], offset[1], offset[2]));
    fromVoxels(rhs, grid);
}
// applyOldPressure, :946-997 (the solution grid starts as 0, :398)
void ref_node_old_pressure(const float *pressure, const int32_t *material, const int64_t res[3], const int32_t *expLabels, const int64_t expRes[3], const int64_t offset[3],
			   double *solution)
{
    Node node(nullptr);
    SIM_RawIndexField mat;
    indexField(mat, material, res);
    SIM_RawField pr;
    cellField(pr, pressure, res);
    UT_VoxelArray<int> labels;
    toVoxels(labels, expLabels, expRes);
    UT_VoxelArray<double> grid;
    grid.size(int(expRes[0]), int(expRes[1]), int(expRes[2]));
    grid.constant(0);
    node.applyOldPressure(grid, pr, mat, labels, UT_Vector3I(offset[0], offset[1], offset[2]));
    fromVoxels(solution, grid);
}
// applySolutionToPressure, :1000-1047: pressure in / out (only LIQUID cells are written)
void ref_node_solution_to_pressure(float *pressure, const int32_t *material, const double *solution, const int64_t res[3], const int32_t *expLabels,
				   const int64_t expRes[3], const int64_t offset[3])
{
    Node node(nullptr);
    SIM_RawIndexField mat;
    indexField(mat, material, res);
    SIM_RawField pr;
    cellField(pr, pressure, res);
    UT_VoxelArray<int> labels;
    toVoxels(labels, expLabels, expRes);
    UT_VoxelArray<double> grid;
    toVoxels(grid, solution, expRes);
    node.applySolutionToPressure(pr, mat, labels, grid, UT_Vector3I(offset[0], offset[1], offset[2]));
    fromVoxels(pressure, *pr.field());
}
// applyPressureGradient, :1050-1131, one axis: velocity in / out
void ref_node_pressure_gradient(float *velocity, const float *cutCell, const float *liquidSurface, const float *pressure, const float *validFaces, const int32_t *material,
				const int64_t res[3], int axis)
{
    Node node(nullptr);
    SIM_RawField vel, cut, liquid, pr, valid;
    faceField(vel, velocity, res, axis);
    faceField(cut, cutCell, res, axis);
    faceField(valid, validFaces, res, axis);
    cellField(liquid, liquidSurface, res);
    cellField(pr, pressure, res);
    SIM_RawIndexField mat;
    indexField(mat, material, res);
    node.applyPressureGradient(vel, cut, liquid, pr, valid, mat, axis);
    fromVoxels(velocity, *vel.field());
}
// The whole node: solveGasSubclass, :113-714 -- fields in, one pressure projection (production wiring: tiled Gauss-Seidel V-cycle as the
// preconditioner, :463-466, or the diagonal one), pressure / velocity / validFaces out.  solidSurface, solidVelocity, pressure may be
// null (the node then takes "no solid", no solid motion, a local zero pressure).  Returns 1 if the node reported success; its log goes
// to `log` (truncated to logCap).
int ref_node_solve(const float *liquidSurface, const float *solidSurface, float *const velocity[3], const float *const cutCell[3], const float *const solidVelocity[3],
		   float *pressure, float *const validFaces[3], float density, const int64_t res[3], double tolerance, int maxIterations, int useMGPreconditioner,
		   int useOldPressure, char *log, int logCap)
{
    Node node(nullptr);
    node.options[SIM_NAME_TOLERANCE] = tolerance;
    node.options["maxIterations"] = maxIterations;
    node.options["useMGPreconditioner"] = useMGPreconditioner;
    node.options["useOldPressure"] = useOldPressure;
    SIM_ScalarField surface, solid, pr, dens;
    SIM_VectorField vel, cut, solidVel, valid;
    cellField(*surface.getField(), liquidSurface, res);
    cellField(*dens.getField(), nullptr, res);
    dens.getField()->makeConstant(density);
    vectorField(vel, velocity, res);
    vectorField(cut, cutCell, res);
    vectorField(valid, nullptr, res);
    SIM_Object obj;
    obj.scalarFields[GAS_NAME_SURFACE] = &surface;
    obj.scalarFields[GAS_NAME_DENSITY] = &dens;
    obj.vectorFields[GAS_NAME_VELOCITY] = &vel;
    obj.vectorFields["cutCellWeights"] = &cut;
    obj.vectorFields["validFaces"] = &valid;
    if (solidSurface) { cellField(*solid.getField(), solidSurface, res); obj.scalarFields[GAS_NAME_COLLISION] = &solid; }
    if (solidVelocity) { vectorField(solidVel, solidVelocity, res); obj.vectorFields[GAS_NAME_COLLISIONVELOCITY] = &solidVel; }
    cellField(*pr.getField(), pressure, res);
    obj.scalarFields[GAS_NAME_PRESSURE] = &pr;
    SIM_Engine engine;
    std::ostringstream captured;
    std::streambuf *old = std::cout.rdbuf(captured.rdbuf());
    const bool ok = node.solveGasSubclass(engine, &obj, 0, 1. / 24.);
    std::cout.rdbuf(old);
    std::string text = captured.str();
    for (const std::string &e : obj.errors) text += "ERROR: " + e + "\n";
    if (log && logCap > 0)
    {
	const size_t n = std::min(text.size(), size_t(logCap - 1));
	std::memcpy(log, text.data() + (text.size() - n), n);  // the tail holds the iteration count and the divergence check
	log[n] = 0;
    }
    fromVoxels(pressure, *pr.getField()->field());
    for (int a = 0; a < 3; ++a)
    {
	fromVoxels(velocity[a], *vel.getField(a)->field());
	fromVoxels(validFaces[a], *valid.getField(a)->field());
    }
    return ok ? 1 : 0;
}
// bench.py --impl reference: use every host core even when the launcher (torchrun) exported OMP_NUM_THREADS=1
void ref_set_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
}
} // extern "C"
