// test_facade_node.cpp -- TEST: the node-level C++ facade (include/gmg_b200_hdk_node.hpp, namespace HDKB200 here: the steps either side of
// the solve with the reference's names and signatures over the C ABI) against the reference's OWN sources -- HDK::Utilities (HDK_Utilities.cpp)
// and the private builders of HDK_GeometricFreeSurfacePressureSolver (the node's own .cpp), compiled unmodified over oracle/shim -- on the same
// SIM fields, in one process.
//
//   built by oracle/Makefile (target ref) into oracle/_ref/test_facade_node, only where the reference sources exist; the binary travels to the
//   GPU box, where tests/test_zz_facade_node.py runs it.   usage: test_facade_node [gridSize = 24]
//
// Gates: material labels, valid faces, domain labels, old pressure, pressure and velocity (integer / fpreal32) bit for bit; boundary weights
// and right-hand sides (fp64, the kernels may contract a multiply-add) <= 1e-14 relative L-inf; one whole projection against the node's own
// solveGasSubclass: valid faces bit for bit, iteration count +-1, pressure <= 1e-5, velocity <= 2e-5 (the north_star's bars).
#include <cstdio>
#include <iostream>
#include <random>
#include <sstream>
#include <string>

#include "HDK_GeometricCGPoissonSolver.h"
#include "HDK_GeometricMultigridOperators.h"
#include "HDK_GeometricMultigridPoissonSolver.h"
#include "HDK_Utilities.h"
#include "hdk_node_shim.h"
// the node's builders are private members: this test reads the class declaration with every member public (see oracle/ref_bridge.cpp)
#define private public
#define protected public
#include "HDK_GeometricFreeSurfacePressureSolver.h"
#undef protected
#undef private

#define GMG_HDK_NAMESPACE HDKB200
#include "gmg_b200_hdk_node.hpp"

namespace Ops = HDK::GeometricMultigridOperators;
namespace New = HDKB200::FreeSurfacePressure;
using Node = HDK_GeometricFreeSurfacePressureSolver;

static int g_fail = 0;
#define EXPECT(cond, ...)                                   \
    do                                                      \
    {                                                       \
	if (!(cond))                                        \
	{                                                   \
	    ++g_fail;                                       \
	    std::printf("FAIL %s:%d  ", __FILE__, __LINE__); \
	    std::printf(__VA_ARGS__);                       \
	    std::printf("\n");                              \
	}                                                   \
    } while (0)

template <typename T>
static long differing(const UT_VoxelArray<T> &a, const UT_VoxelArray<T> &b)
{
    if (!(a.getVoxelRes() == b.getVoxelRes())) return -1;
    const UT_Vector3I r = a.getVoxelRes();
    long n = 0;
    for (int z = 0; z < r[2]; ++z)
	for (int y = 0; y < r[1]; ++y)
	    for (int x = 0; x < r[0]; ++x)
		if (!(a(x, y, z) == b(x, y, z))) ++n;
    return n;
}
static double relDiff(const UT_VoxelArray<double> &a, const UT_VoxelArray<double> &b)
{
    const UT_Vector3I r = a.getVoxelRes();
    double d = 0, s = 0;
    for (int z = 0; z < r[2]; ++z)
	for (int y = 0; y < r[1]; ++y)
	    for (int x = 0; x < r[0]; ++x)
	    {
		d = std::max(d, std::fabs(a(x, y, z) - b(x, y, z)));
		s = std::max(s, std::fabs(b(x, y, z)));
	    }
    return d / std::max(s, 1e-300);
}

struct Fields
{
    int n;
    SIM_RawField liquid, solid, pressure;
    SIM_VectorField cut, velocity, solidVelocityAligned, solidVelocityShifted;
};

// a tank with solid walls (closed faces), a tilted free surface, open / fractional cut-cell weights, random velocities and pressure;
// one solid-velocity field on the velocity's own lattice, one on a lattice shifted by a third of a cell (so it has to be sampled)
static void makeFields(Fields &f, int n, unsigned seed)
{
    f.n = n;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> u(0.f, 1.f);
    f.liquid.init(n, n, n);
    f.solid.init(n, n, n);
    f.pressure.init(n, n, n);
    f.cut.initFaces(n, n, n);
    f.velocity.initFaces(n, n, n);
    f.solidVelocityAligned.initFaces(n, n, n);
    auto wall = [n](int x, int y, int z) { return x == 0 || x == n - 1 || z == 0 || z == n - 1 || y == 0; };
    for (int z = 0; z < n; ++z)
	for (int y = 0; y < n; ++y)
	    for (int x = 0; x < n; ++x)
	    {
		f.liquid.fieldNC()->setValue(x, y, z, ((y - 0.55f * n) + 0.15f * (x - n / 2.f) + 0.1f * (z - n / 2.f)) / n);
		f.solid.fieldNC()->setValue(x, y, z, u(rng) - 0.5f);  // both signs: both branches of isCellLiquid
		f.pressure.fieldNC()->setValue(x, y, z, u(rng));
	    }
    for (int a = 0; a < 3; ++a)
    {
	const UT_Vector3I r = f.cut.getField(a)->field()->getVoxelRes();
	for (int z = 0; z < r[2]; ++z)
	    for (int y = 0; y < r[1]; ++y)
		for (int x = 0; x < r[0]; ++x)
		{
		    int b[3] = {x, y, z}, c[3] = {x, y, z};
		    --b[a];
		    const bool inRange = b[a] >= 0 && c[a] < n;
		    float w = 0.f;
		    if (inRange && !wall(b[0], b[1], b[2]) && !wall(c[0], c[1], c[2])) w = u(rng) < 0.3f ? 0.05f + 0.9f * u(rng) : 1.f;
		    f.cut.getField(a)->fieldNC()->setValue(x, y, z, w);
		    f.velocity.getField(a)->fieldNC()->setValue(x, y, z, w > 0 ? u(rng) - 0.5f : 0.f);
		    f.solidVelocityAligned.getField(a)->fieldNC()->setValue(x, y, z, 0.25f * u(rng));
		}
    }
    // same resolution, origin moved by a third of a cell: not aligned with the velocity field
    for (int a = 0; a < 3; ++a)
    {
	SIM_RawField *g = f.solidVelocityShifted.getField(a);
	g->init(SIM_FieldSample(SIM_SAMPLE_FACEX + a), UT_Vector3(1.f / 3, 1.f / 3, 1.f / 3), UT_Vector3(float(n), float(n), float(n)), n, n, n);
	const UT_Vector3I r = g->field()->getVoxelRes();
	for (int z = 0; z < r[2]; ++z)
	    for (int y = 0; y < r[1]; ++y)
		for (int x = 0; x < r[0]; ++x) g->fieldNC()->setValue(x, y, z, 0.5f * u(rng));
    }
}

template <typename T>
static void copyGrid(UT_VoxelArray<T> &dst, const UT_VoxelArray<T> &src)
{
    const UT_Vector3I r = src.getVoxelRes();
    for (int z = 0; z < r[2]; ++z)
	for (int y = 0; y < r[1]; ++y)
	    for (int x = 0; x < r[0]; ++x) dst.setValue(x, y, z, src(x, y, z));
}
static void copyVector(SIM_VectorField &dst, const SIM_VectorField &src, int n)
{
    dst.initFaces(n, n, n);
    for (int a = 0; a < 3; ++a) copyGrid(*dst.getField(a)->fieldNC(), *src.getField(a)->field());
}
static double maxAbs(const UT_VoxelArray<fpreal32> &a)
{
    const UT_Vector3I r = a.getVoxelRes();
    double m = 0;
    for (int z = 0; z < r[2]; ++z)
	for (int y = 0; y < r[1]; ++y)
	    for (int x = 0; x < r[0]; ++x) m = std::max(m, double(std::fabs(a(x, y, z))));
    return m;
}
static double maxDiff(const UT_VoxelArray<fpreal32> &a, const UT_VoxelArray<fpreal32> &b)
{
    const UT_Vector3I r = a.getVoxelRes();
    double m = 0;
    for (int z = 0; z < r[2]; ++z)
	for (int y = 0; y < r[1]; ++y)
	    for (int x = 0; x < r[0]; ++x) m = std::max(m, double(std::fabs(a(x, y, z) - b(x, y, z))));
    return m;
}

// One whole pressure projection: the reference node's own solveGasSubclass (HDK_GeometricFreeSurfacePressureSolver.cpp:113-714, unmodified,
// production wiring: tiled Gauss-Seidel V-cycle inside the PCG) against the same body written with the facade -- what the patched node runs.
static void wholeProjection(const Fields &f, bool withSolid)
{
    const int n = f.n;
    const double tolerance = 1e-7;
    const int maxIterations = 400;
    // ---- the reference node on its own SIM_Object
    SIM_ScalarField surface, collision, density, pressureR;
    SIM_VectorField velocityR, cutR, validR, collisionVel;
    surface.getField()->init(n, n, n);
    copyGrid(*surface.getField()->fieldNC(), *f.liquid.field());
    density.getField()->init(n, n, n);
    density.getField()->makeConstant(1000.f);
    pressureR.getField()->init(n, n, n);
    copyVector(velocityR, f.velocity, n);
    copyVector(cutR, f.cut, n);
    validR.initFaces(n, n, n);
    SIM_Object obj;
    obj.scalarFields[GAS_NAME_SURFACE] = &surface;
    obj.scalarFields[GAS_NAME_DENSITY] = &density;
    obj.scalarFields[GAS_NAME_PRESSURE] = &pressureR;
    obj.vectorFields[GAS_NAME_VELOCITY] = &velocityR;
    obj.vectorFields["cutCellWeights"] = &cutR;
    obj.vectorFields["validFaces"] = &validR;
    if (withSolid)
    {
	collision.getField()->init(n, n, n);
	copyGrid(*collision.getField()->fieldNC(), *f.solid.field());
	copyVector(collisionVel, f.solidVelocityAligned, n);
	obj.scalarFields[GAS_NAME_COLLISION] = &collision;
	obj.vectorFields[GAS_NAME_COLLISIONVELOCITY] = &collisionVel;
    }
    Node node(nullptr);
    node.options[SIM_NAME_TOLERANCE] = tolerance;
    node.options["maxIterations"] = maxIterations;
    node.options["useMGPreconditioner"] = 1;
    node.options["useOldPressure"] = 0;
    SIM_Engine engine;
    std::ostringstream captured;
    std::streambuf *old = std::cout.rdbuf(captured.rdbuf());
    const bool ok = node.solveGasSubclass(engine, &obj, 0, 1. / 24.);
    std::cout.rdbuf(old);
    EXPECT(ok && obj.errors.empty(), "the reference node failed: %s", obj.errors.empty() ? "?" : obj.errors[0].c_str());
    int iterationsR = -1;
    {
	const std::string text = captured.str();
	const size_t at = text.rfind("Iterations: ");
	if (at != std::string::npos) iterationsR = std::atoi(text.c_str() + at + 12);
    }

    // ---- the same body with the facade (the calls of solveGasSubclass, GFS.cpp:262-660, re-qualified)
    namespace NewOps = HDKB200::GeometricMultigridOperators;
    const std::array<const SIM_RawField *, 3> cut = {f.cut.getField(0), f.cut.getField(1), f.cut.getField(2)};
    SIM_RawField noSolid;
    noSolid.init(n, n, n);
    noSolid.makeConstant(-10.f);  // GFS.cpp:207-219: no collision field = "all fluid", -10 dx
    const SIM_RawField &solidSurface = withSolid ? f.solid : noSolid;
    SIM_RawIndexField material;
    HDKB200::Utilities::buildMaterialCellLabels(material, f.liquid, solidSurface, cut);
    SIM_VectorField valid, velocity;
    valid.initFaces(n, n, n);
    copyVector(velocity, f.velocity, n);
    New::buildValidFaces(valid, material, cut);
    for (int a = 0; a < 3; ++a)
	EXPECT(differing(*valid.getField(a)->field(), *validR.getField(a)->field()) == 0, "projection: valid faces differ on axis %d", a);
    UT_VoxelArray<int> base;
    base.size(n, n, n);
    base.constant(NewOps::EXTERIOR_CELL);
    New::buildMGDomainLabels(base, material);
    std::array<UT_VoxelArray<double>, 3> baseW;
    for (int a = 0; a < 3; ++a)
    {
	UT_Vector3I size(n, n, n);
	++size[a];
	baseW[a].size(int(size[0]), int(size[1]), int(size[2]));
	baseW[a].constant(0);
	New::buildMGBoundaryWeights(baseW[a], *cut[a], f.liquid, *valid.getField(a), material, base, a);
    }
    UT_VoxelArray<int> labels;
    auto isExt = [](const int v) { return v == NewOps::EXTERIOR_CELL; };
    auto isInt = [](const int v) { return v == NewOps::INTERIOR_CELL; };
    auto isDir = [](const int v) { return v == NewOps::DIRICHLET_CELL; };
    const std::pair<UT_Vector3I, int> settings = NewOps::buildExpandedCellLabels(labels, base, isExt, isInt, isDir);
    const UT_Vector3I offset = settings.first;
    const int mgLevels = settings.second;
    std::array<UT_VoxelArray<double>, 3> weights;
    for (int a = 0; a < 3; ++a)
    {
	UT_Vector3I size = labels.getVoxelRes();
	++size[a];
	weights[a].size(int(size[0]), int(size[1]), int(size[2]));
	weights[a].constant(0);
	NewOps::buildExpandedBoundaryWeights(weights[a], baseW[a], labels, offset, a);
    }
    NewOps::setBoundaryCellLabels(labels, weights);
    const UT_Vector3I er = labels.getVoxelRes();
    UT_VoxelArray<double> rhs, solution;
    rhs.size(int(er[0]), int(er[1]), int(er[2]));
    rhs.constant(0);
    solution.size(int(er[0]), int(er[1]), int(er[2]));
    solution.constant(0);
    New::buildRHS(rhs, material, velocity, withSolid ? &f.solidVelocityAligned : nullptr, cut, labels, offset);
    HDKB200::GeometricMultigridPoissonSolver mg(labels, weights, mgLevels, true /* useGaussSeidel, GFS.cpp:463-466 */);
    std::vector<double> history;
    const int iterations = HDKB200::solveGeometricConjugateGradient(mg, solution, rhs, tolerance, maxIterations, true, &history);
    SIM_RawField pressure;
    pressure.init(n, n, n);
    pressure.makeConstant(0);
    New::applySolutionToPressure(pressure, material, labels, solution, offset);
    for (int a = 0; a < 3; ++a) New::applyPressureGradient(*velocity.getField(a), *cut[a], f.liquid, pressure, *valid.getField(a), material, a);

    EXPECT(iterationsR >= 0 && std::abs(iterations - iterationsR) <= 1, "projection: iterations %d vs the node's %d", iterations, iterationsR);
    const double pScale = maxAbs(*pressureR.getField()->field());
    const double pDiff = maxDiff(*pressure.field(), *pressureR.getField()->field());
    EXPECT(pScale > 0 && pDiff <= 1e-5 * pScale, "projection: pressure deviates by %.3e of %.3e", pDiff, pScale);
    for (int a = 0; a < 3; ++a)
    {
	const double vDiff = maxDiff(*velocity.getField(a)->field(), *velocityR.getField(a)->field());
	EXPECT(vDiff <= 2e-5 * std::max(pScale, maxAbs(*velocityR.getField(a)->field())), "projection: velocity axis %d deviates by %.3e", a, vDiff);
    }
    std::printf("projection (%s): %d iterations (node: %d), pressure deviation %.2e of %.2e\n", withSolid ? "solid field, moving solid" : "no solid", iterations,
		iterationsR, pDiff, pScale);
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? std::atoi(argv[1]) : 24;
    Fields f;
    makeFields(f, n, 7u + unsigned(n));
    const std::array<const SIM_RawField *, 3> cut = {f.cut.getField(0), f.cut.getField(1), f.cut.getField(2)};
    Node node(nullptr);
    using MGCellLabels = Ops::CellLabels;

    // ---- material labels (HDK_Utilities.cpp:87-148)
    SIM_RawIndexField matR, matN;
    matR.match(f.liquid);
    HDK::Utilities::buildMaterialCellLabels(matR, f.liquid, f.solid, cut);
    HDKB200::Utilities::buildMaterialCellLabels(matN, f.liquid, f.solid, cut);
    EXPECT(differing(*matR.field(), *matN.field()) == 0, "material labels differ in %ld cells", differing(*matR.field(), *matN.field()));
    long liquidCells = 0;
    for (int z = 0; z < n; ++z)
	for (int y = 0; y < n; ++y)
	    for (int x = 0; x < n; ++x) liquidCells += (*matR.field())(x, y, z) == HDK::Utilities::LIQUID_CELL;
    EXPECT(liquidCells > n * n, "degenerate test fields: %ld liquid cells", liquidCells);

    // ---- valid faces (GFS.cpp:717-744)
    SIM_VectorField validR, validN;
    validR.initFaces(n, n, n);
    validN.initFaces(n, n, n);
    node.buildValidFaces(validR, matR, cut);
    New::buildValidFaces(validN, matR, cut);
    for (int a = 0; a < 3; ++a)
	EXPECT(differing(*validR.getField(a)->field(), *validN.getField(a)->field()) == 0, "valid faces differ on axis %d", a);

    // ---- domain labels (GFS.cpp:746-793)
    UT_VoxelArray<int> baseR, baseN;
    baseR.size(n, n, n);
    baseR.constant(MGCellLabels::EXTERIOR_CELL);
    baseN.size(n, n, n);
    baseN.constant(MGCellLabels::EXTERIOR_CELL);
    node.buildMGDomainLabels(baseR, matR);
    New::buildMGDomainLabels(baseN, matR);
    EXPECT(differing(baseR, baseN) == 0, "domain labels differ");

    // ---- boundary weights (GFS.cpp:796-865)
    std::array<UT_VoxelArray<double>, 3> wR, wN;
    for (int a = 0; a < 3; ++a)
    {
	UT_Vector3I size(n, n, n);
	++size[a];
	wR[a].size(int(size[0]), int(size[1]), int(size[2]));
	wR[a].constant(0);
	wN[a].size(int(size[0]), int(size[1]), int(size[2]));
	wN[a].constant(0);
	node.buildMGBoundaryWeights(wR[a], *cut[a], f.liquid, *validR.getField(a), matR, baseR, a);
	New::buildMGBoundaryWeights(wN[a], *cut[a], f.liquid, *validR.getField(a), matR, baseR, a);
	EXPECT(relDiff(wN[a], wR[a]) <= 1e-14, "boundary weights, axis %d: %.3e", a, relDiff(wN[a], wR[a]));
    }

    // ---- expanded domain with the reference's own builders (GFS.cpp:333-365)
    UT_VoxelArray<int> labels;
    auto isExt = [](const int v) { return v == MGCellLabels::EXTERIOR_CELL; };
    auto isInt = [](const int v) { return v == MGCellLabels::INTERIOR_CELL; };
    auto isDir = [](const int v) { return v == MGCellLabels::DIRICHLET_CELL; };
    const std::pair<UT_Vector3I, int> settings = Ops::buildExpandedCellLabels(labels, baseR, isExt, isInt, isDir);
    const UT_Vector3I offset = settings.first;
    std::array<UT_VoxelArray<double>, 3> weights;
    for (int a = 0; a < 3; ++a)
    {
	UT_Vector3I size = labels.getVoxelRes();
	++size[a];
	weights[a].size(int(size[0]), int(size[1]), int(size[2]));
	weights[a].constant(0);
	Ops::buildExpandedBoundaryWeights(weights[a], wR[a], labels, offset, a);
    }
    Ops::setBoundaryCellLabels(labels, weights);
    const UT_Vector3I er = labels.getVoxelRes();

    // ---- right-hand side (GFS.cpp:868-943): no solid motion, an aligned solid-velocity field, one that has to be sampled
    const SIM_VectorField *solids[3] = {nullptr, &f.solidVelocityAligned, &f.solidVelocityShifted};
    for (int k = 0; k < 3; ++k)
    {
	UT_VoxelArray<double> rhsR, rhsN;
	rhsR.size(int(er[0]), int(er[1]), int(er[2]));
	rhsR.constant(0);
	rhsN.size(int(er[0]), int(er[1]), int(er[2]));
	rhsN.constant(0);
	node.buildRHS(rhsR, matR, f.velocity, solids[k], cut, labels, offset);
	New::buildRHS(rhsN, matR, f.velocity, solids[k], cut, labels, offset);
	EXPECT(relDiff(rhsN, rhsR) <= 1e-14, "right-hand side, solid velocity variant %d: %.3e", k, relDiff(rhsN, rhsR));
    }

    // ---- warm start (GFS.cpp:946-997)
    UT_VoxelArray<double> solR, solN;
    solR.size(int(er[0]), int(er[1]), int(er[2]));
    solR.constant(0);
    solN.size(int(er[0]), int(er[1]), int(er[2]));
    solN.constant(0);
    node.applyOldPressure(solR, f.pressure, matR, labels, offset);
    New::applyOldPressure(solN, f.pressure, matR, labels, offset);
    EXPECT(differing(solR, solN) == 0, "applyOldPressure differs");

    // ---- solution -> pressure (GFS.cpp:1000-1047) on a random solution grid
    {
	std::mt19937_64 rng(99);
	std::uniform_real_distribution<double> u(-1, 1);
	UT_VoxelArray<double> x;
	x.size(int(er[0]), int(er[1]), int(er[2]));
	x.constant(0);
	for (int z = 0; z < er[2]; ++z)
	    for (int y = 0; y < er[1]; ++y)
		for (int xx = 0; xx < er[0]; ++xx)
		    if (labels(xx, y, z) == MGCellLabels::INTERIOR_CELL || labels(xx, y, z) == MGCellLabels::BOUNDARY_CELL) x.setValue(xx, y, z, u(rng));
	SIM_RawField pR, pN;
	pR.init(n, n, n);
	pR.makeConstant(-1.f);
	pN.init(n, n, n);
	pN.makeConstant(-1.f);
	node.applySolutionToPressure(pR, matR, labels, x, offset);
	New::applySolutionToPressure(pN, matR, labels, x, offset);
	EXPECT(differing(*pR.field(), *pN.field()) == 0, "applySolutionToPressure differs");
    }

    // ---- velocity update (GFS.cpp:1050-1131)
    for (int a = 0; a < 3; ++a)
    {
	SIM_RawField vR, vN;
	vR.match(*f.velocity.getField(a));
	vN.match(*f.velocity.getField(a));
	const UT_Vector3I r = vR.field()->getVoxelRes();
	for (int z = 0; z < r[2]; ++z)
	    for (int y = 0; y < r[1]; ++y)
		for (int x = 0; x < r[0]; ++x)
		{
		    vR.fieldNC()->setValue(x, y, z, (*f.velocity.getField(a)->field())(x, y, z));
		    vN.fieldNC()->setValue(x, y, z, (*f.velocity.getField(a)->field())(x, y, z));
		}
	node.applyPressureGradient(vR, *cut[a], f.liquid, f.pressure, *validR.getField(a), matR, a);
	New::applyPressureGradient(vN, *cut[a], f.liquid, f.pressure, *validR.getField(a), matR, a);
	EXPECT(differing(*vR.field(), *vN.field()) == 0, "applyPressureGradient differs on axis %d", a);
    }

    // ---- one whole projection, the node against the facade
    wholeProjection(f, false);
    wholeProjection(f, true);

    if (g_fail) { std::printf("FACADE_NODE_FAILED %d\n", g_fail); return 1; }
    std::printf("FACADE_NODE_OK gridSize %d, %ld liquid cells\n", n, liquidCells);
    return 0;
}
