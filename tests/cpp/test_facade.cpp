// test_facade.cpp -- TEST: the C++ host facade (include/gmg_b200_hdk.hpp, namespace HDKB200 here) against
// the reference's OWN sources (namespace HDK, compiled unmodified from /root/reference/Source over
// oracle/shim) on the same UT_VoxelArray inputs, in one process.
//
//   built by oracle/Makefile (target ref) into oracle/_ref/test_facade -- only where the reference
//   sources exist; the binary travels to the GPU box, where tests/test_gpu_facade.py runs it.
//   usage: test_facade <input.bin>     (input written by the pytest wrapper from domains.py)
//
// Gates: labels / coarse labels / boundary lists bit-exact; every operator <= 1e-13 relative L-inf;
// PCG residual history <= 1e-5 relative per iteration, iteration count +-1, pressure <= 1e-5 relative L-inf.
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>

#include "HDK_GeometricCGPoissonSolver.h"
#include "HDK_GeometricMultigridOperators.h"
#include "HDK_GeometricMultigridPoissonSolver.h"

#define GMG_HDK_NAMESPACE HDKB200
#include "gmg_b200_hdk.hpp"

namespace Ref = HDK::GeometricMultigridOperators;
namespace New = HDKB200::GeometricMultigridOperators;
using Weights = std::array<UT_VoxelArray<double>, 3>;

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
static bool sameGrid(const UT_VoxelArray<T> &a, const UT_VoxelArray<T> &b)
{
    if (!(a.getVoxelRes() == b.getVoxelRes())) return false;
    const UT_Vector3I r = a.getVoxelRes();
    for (int z = 0; z < r[2]; ++z)
	for (int y = 0; y < r[1]; ++y)
	    for (int x = 0; x < r[0]; ++x)
		if (!(a(x, y, z) == b(x, y, z))) return false;
    return true;
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

static void randomActive(UT_VoxelArray<double> &v, const UT_VoxelArray<int> &labels, unsigned seed)
{
    const UT_Vector3I r = labels.getVoxelRes();
    v.size(int(r[0]), int(r[1]), int(r[2]));
    v.constant(0);
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u(-1, 1);
    for (int z = 0; z < r[2]; ++z)
	for (int y = 0; y < r[1]; ++y)
	    for (int x = 0; x < r[0]; ++x)
	    {
		const int l = labels(x, y, z);
		if (l == Ref::INTERIOR_CELL || l == Ref::BOUNDARY_CELL) v.setValue(x, y, z, u(rng));
	    }
}

static std::vector<double> parseHistory(const std::string &out, int *iterations)
{
    std::vector<double> h;
    std::istringstream in(out);
    std::string line;
    *iterations = -1;
    while (std::getline(in, line))
    {
	const auto p = line.find("Relative error:");
	if (p != std::string::npos) h.push_back(std::atof(line.c_str() + p + 15));
	const auto q = line.find("Iterations:");
	if (q != std::string::npos) *iterations = std::atoi(line.c_str() + q + 11);
    }
    return h;
}

static void onSegv(int sig)
{
    void *frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "test_facade: fatal signal, backtrace:\n";
    if (write(2, msg, sizeof(msg) - 1) < 0) {}
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

int main(int argc, char **argv)
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    signal(SIGSEGV, onSegv);
    signal(SIGBUS, onSegv);
    if (argc < 2) { std::printf("usage: test_facade <input.bin>\n"); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    int32_t n[3];
    in.read(reinterpret_cast<char *>(n), sizeof(n));
    UT_VoxelArray<int> baseLabels;
    baseLabels.size(n[0], n[1], n[2]);
    {
	std::vector<int32_t> buf(size_t(n[0]) * n[1] * n[2]);
	in.read(reinterpret_cast<char *>(buf.data()), buf.size() * 4);
	for (int z = 0; z < n[2]; ++z)
	    for (int y = 0; y < n[1]; ++y)
		for (int x = 0; x < n[0]; ++x) baseLabels.setValue(x, y, z, buf[x + size_t(n[0]) * (y + size_t(n[1]) * z)]);
    }
    Weights baseW;
    for (int a = 0; a < 3; ++a)
    {
	int r[3] = {n[0], n[1], n[2]};
	r[a] += 1;
	baseW[a].size(r[0], r[1], r[2]);
	std::vector<double> buf(size_t(r[0]) * r[1] * r[2]);
	in.read(reinterpret_cast<char *>(buf.data()), buf.size() * 8);
	for (int z = 0; z < r[2]; ++z)
	    for (int y = 0; y < r[1]; ++y)
		for (int x = 0; x < r[0]; ++x) baseW[a].setValue(x, y, z, buf[x + size_t(r[0]) * (y + size_t(r[1]) * z)]);
    }
    double dx = 0;
    in.read(reinterpret_cast<char *>(&dx), 8);
    if (!in) { std::printf("short input file\n"); return 2; }

    auto isExt = [](int l) { return l == Ref::EXTERIOR_CELL; };
    auto isInt = [](int l) { return l == Ref::INTERIOR_CELL; };
    auto isDir = [](int l) { return l == Ref::DIRICHLET_CELL; };

    // ---- domain builders (the sequence of Test.cpp:170-204 / GFS.cpp:344-365) -------------------------------------
    UT_VoxelArray<int> labR, labN;
    auto pr = Ref::buildExpandedCellLabels(labR, baseLabels, isExt, isInt, isDir);
    auto pn = New::buildExpandedCellLabels(labN, baseLabels, isExt, isInt, isDir);
    EXPECT(pr.first == pn.first && pr.second == pn.second, "offset / mgLevels differ");
    EXPECT(sameGrid(labR, labN), "buildExpandedCellLabels differs");
    Weights wR, wN;
    for (int a = 0; a < 3; ++a)
    {
	// the reference expects the caller to size the expanded face grid (Test.cpp:182-197 does); the facade sizes its own
	UT_Vector3I fr = labR.getVoxelRes();
	fr[a] += 1;
	wR[a].size(int(fr[0]), int(fr[1]), int(fr[2]));
	wR[a].constant(0);
	Ref::buildExpandedBoundaryWeights(wR[a], baseW[a], labR, pr.first, a);
	New::buildExpandedBoundaryWeights(wN[a], baseW[a], labN, pn.first, a);
	EXPECT(sameGrid(wR[a], wN[a]), "buildExpandedBoundaryWeights differs on axis %d", a);
    }
    Ref::setBoundaryCellLabels(labR, wR);
    New::setBoundaryCellLabels(labN, wN);
    EXPECT(sameGrid(labR, labN), "setBoundaryCellLabels differs");
    EXPECT(Ref::unitTestBoundaryCells(labN, &wN), "reference invariant checker rejects our labels (Ops.h:1771-1870)");
    EXPECT(Ref::unitTestExteriorCells(labN), "reference exterior checker rejects our labels (Ops.cpp:602-632)");
    const int mgLevels = pr.second;

    UT_VoxelArray<int> coarseR = Ref::buildCoarseCellLabels(labR), coarseN = New::buildCoarseCellLabels(labN);
    EXPECT(sameGrid(coarseR, coarseN), "buildCoarseCellLabels differs");
    EXPECT(Ref::unitTestCoarsening(coarseN, labN), "reference coarsening checker rejects our coarse labels (Ops.cpp:471-600)");
    UT_Array<UT_Vector3I> bandR = Ref::buildBoundaryCells(labR, 3), bandN = New::buildBoundaryCells(labN, 3);
    EXPECT(bandR.size() == bandN.size(), "buildBoundaryCells count %d vs %d", int(bandR.size()), int(bandN.size()));
    if (bandR.size() == bandN.size())
    {
	bool same = true;
	for (exint i = 0; i < bandR.size(); ++i) same = same && (bandR[i] == bandN[i]);
	EXPECT(same, "buildBoundaryCells order/content differs");
    }
    std::printf("builders: labels, weights, coarse labels, %d boundary cells compared\n", int(bandR.size()));

    // ---- stateless operators -----------------------------------------------------------------------------------------
    UT_VoxelArray<double> x, b;
    randomActive(x, labR, 1);
    randomActive(b, labR, 2);
    const double tol = 1e-13;
    {
	UT_VoxelArray<double> r = x, q = x;
	Ref::applyPoissonMatrix<double>(r, x, labR, &wR);
	New::applyPoissonMatrix<double>(q, x, labN, &wN);
	EXPECT(relDiff(q, r) <= tol, "applyPoissonMatrix %.3e", relDiff(q, r));
	Ref::computePoissonResidual<double>(r, x, b, labR, &wR);
	New::computePoissonResidual<double>(q, x, b, labN, &wN);
	EXPECT(relDiff(q, r) <= tol, "computePoissonResidual %.3e", relDiff(q, r));
	r = x; q = x;
	Ref::jacobiPoissonSmoother<double>(r, b, labR, &wR);
	New::jacobiPoissonSmoother<double>(q, b, labN, &wN);
	EXPECT(relDiff(q, r) <= tol, "jacobiPoissonSmoother %.3e", relDiff(q, r));
	r = x; q = x;
	Ref::boundaryJacobiPoissonSmoother<double>(r, b, labR, bandR, &wR);
	New::boundaryJacobiPoissonSmoother<double>(q, b, labN, bandN, &wN);
	EXPECT(relDiff(q, r) <= tol, "boundaryJacobiPoissonSmoother %.3e", relDiff(q, r));
	// tiled Gauss-Seidel, the production smoother (GFS.cpp:463-466): the four half-passes of MG.cpp:466-479 / :740-751 in sequence
	r = x; q = x;
	const bool gsOdd[4] = {true, false, false, true}, gsFwd[4] = {true, true, false, false};
	for (int p = 0; p < 4; ++p)
	{
	    Ref::tiledGaussSeidelPoissonSmoother<double>(r, b, labR, gsOdd[p], gsFwd[p], &wR);
	    New::tiledGaussSeidelPoissonSmoother<double>(q, b, labN, gsOdd[p], gsFwd[p], &wN);
	    EXPECT(relDiff(q, r) <= tol, "tiledGaussSeidelPoissonSmoother pass %d %.3e", p, relDiff(q, r));
	}
	// no-weights form on the coarse labels (what the V-cycle uses above level 0)
	UT_VoxelArray<double> xc, bc;
	randomActive(xc, coarseR, 3);
	randomActive(bc, coarseR, 4);
	UT_VoxelArray<double> rc = xc, qc = xc;
	Ref::jacobiPoissonSmoother<double>(rc, bc, coarseR);
	New::jacobiPoissonSmoother<double>(qc, bc, coarseN);
	EXPECT(relDiff(qc, rc) <= tol, "jacobiPoissonSmoother (no weights) %.3e", relDiff(qc, rc));
	rc = xc; qc = xc;
	Ref::downsample<double>(rc, x, coarseR, labR);
	New::downsample<double>(qc, x, coarseN, labN);
	EXPECT(relDiff(qc, rc) <= tol, "downsample %.3e", relDiff(qc, rc));
	r = x; q = x;
	Ref::upsampleAndAdd<double>(r, xc, labR, coarseR);
	New::upsampleAndAdd<double>(q, xc, labN, coarseN);
	EXPECT(relDiff(q, r) <= tol, "upsampleAndAdd %.3e", relDiff(q, r));
	const double dR = Ref::dotProduct<double>(x, b, labR), dN = New::dotProduct<double>(x, b, labN);
	EXPECT(std::fabs(dR - dN) <= 1e-12 * std::fabs(dR) + 1e-12, "dotProduct %.17g vs %.17g", dR, dN);
	const double nR = Ref::squaredL2Norm<double>(x, labR), nN = New::squaredL2Norm<double>(x, labN);
	EXPECT(std::fabs(nR - nN) <= 1e-12 * nR, "squaredL2Norm %.17g vs %.17g", nR, nN);
	EXPECT(std::fabs(Ref::l2Norm<double>(x, labR) - New::l2Norm<double>(x, labN)) <= 1e-12 * std::sqrt(nR), "l2Norm");
	EXPECT(Ref::infNorm(x, labR) == New::infNorm(x, labN), "infNorm %.17g vs %.17g", Ref::infNorm(x, labR), New::infNorm(x, labN));
	r = x; q = x;
	Ref::addToVector<double>(r, b, 0.37, labR);
	New::addToVector<double>(q, b, 0.37, labN);
	EXPECT(relDiff(q, r) <= tol, "addToVector %.3e", relDiff(q, r));
	Ref::addVectors<double>(r, x, b, -0.61, labR);
	New::addVectors<double>(q, x, b, -0.61, labN);
	EXPECT(relDiff(q, r) <= tol, "addVectors %.3e", relDiff(q, r));
	r = x; q = x;
	Ref::scaleVector<double>(r, 1.7, labR);
	New::scaleVector<double>(q, 1.7, labN);
	EXPECT(relDiff(q, r) <= tol, "scaleVector %.3e", relDiff(q, r));
	std::printf("operators: 15 compared at %.0e relative\n", tol);
    }

    // ---- solver: V-cycle and PCG (GFS.cpp:426-484 / Test.cpp:796-832, useGaussSeidel = false) ---------------------------
    {
	// rhs: dx^2-scaled random values on active cells (Test.cpp:793-794 scaling)
	UT_VoxelArray<double> rhs;
	randomActive(rhs, labR, 7);
	Ref::scaleVector<double>(rhs, dx * dx, labR);
	std::ostringstream quiet;
	std::streambuf *old = std::cout.rdbuf(quiet.rdbuf());
	HDK::GeometricMultigridPoissonSolver mgR(labR, wR, mgLevels, false);
	std::cout.rdbuf(old);
	HDKB200::GeometricMultigridPoissonSolver mgN(labN, wN, mgLevels, false);
	EXPECT(mgR.getMGLevels() == mgN.getMGLevels(), "getMGLevels %d vs %d", mgR.getMGLevels(), mgN.getMGLevels());
	UT_VoxelArray<double> zR = rhs, zN = rhs;
	zR.constant(0); zN.constant(0);
	mgR.applyVCycle(zR, rhs);
	mgN.applyVCycle(zN, rhs);
	EXPECT(relDiff(zN, zR) <= 1e-11, "applyVCycle %.3e", relDiff(zN, zR));
	UT_VoxelArray<double> gR = x, gN = x;
	mgR.applyVCycle(gR, rhs, true);
	mgN.applyVCycle(gN, rhs, true);
	EXPECT(relDiff(gN, gR) <= 1e-11, "applyVCycle(useInitialGuess) %.3e", relDiff(gN, gR));

	auto A = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &s) { Ref::applyPoissonMatrix<double>(d, s, labR, &wR); };
	auto M = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &s) { mgR.applyVCycle(d, s); };
	auto dot = [&](const UT_VoxelArray<double> &a, const UT_VoxelArray<double> &c) { return Ref::dotProduct<double>(a, c, labR); };
	auto nrm = [&](const UT_VoxelArray<double> &a) { return Ref::squaredL2Norm<double>(a, labR); };
	auto axpy = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &s, double sc) { Ref::addToVector<double>(d, s, sc, labR); };
	auto addS = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &u, const UT_VoxelArray<double> &s, double sc) {
	    Ref::addVectors<double>(d, u, s, sc, labR);
	};
	UT_VoxelArray<double> pR = rhs, pN = rhs;
	pR.constant(0); pN.constant(0);
	std::ostringstream cap;
	old = std::cout.rdbuf(cap.rdbuf());
	HDK::solveGeometricConjugateGradient(pR, rhs, A, M, dot, nrm, axpy, addS, 1e-6, 1000);
	std::cout.rdbuf(old);
	int itR = -1;
	const std::vector<double> hR = parseHistory(cap.str(), &itR);
	std::vector<double> hN;
	const int itN = HDKB200::solveGeometricConjugateGradient(mgN, pN, rhs, 1e-6, 1000, true, &hN);
	EXPECT(std::abs(itR - itN) <= 1, "iterations %d vs %d", itR, itN);
	const size_t m = std::min(hR.size(), hN.size());
	double worst = 0;
	// the reference prints 6 significant digits; compare at that resolution plus the 1e-5 gate
	for (size_t i = 0; i < m; ++i) worst = std::max(worst, std::fabs(hR[i] - hN[i]) / hR[i]);
	EXPECT(m > 0 && worst <= 1e-5 + 6e-6, "residual history deviates by %.3e", worst);
	EXPECT(relDiff(pN, pR) <= 1e-5, "pressure %.3e", relDiff(pN, pR));
	std::printf("solver: V-cycle %.2e, PCG iterations %d vs %d, history dev %.2e (6-digit prints), pressure %.2e\n", relDiff(zN, zR), itR, itN, worst,
		    relDiff(pN, pR));

	// the functor form of the facade, driven by the facade's own stateless operators and solver
	auto A2 = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &s) { New::applyPoissonMatrix<double>(d, s, labN, &wN); };
	auto M2 = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &s) { mgN.applyVCycle(d, s); };
	auto dot2 = [&](const UT_VoxelArray<double> &a, const UT_VoxelArray<double> &c) { return New::dotProduct<double>(a, c, labN); };
	auto nrm2 = [&](const UT_VoxelArray<double> &a) { return New::squaredL2Norm<double>(a, labN); };
	auto axpy2 = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &s, double sc) { New::addToVector<double>(d, s, sc, labN); };
	auto addS2 = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &u, const UT_VoxelArray<double> &s, double sc) {
	    New::addVectors<double>(d, u, s, sc, labN);
	};
	UT_VoxelArray<double> pF = rhs;
	pF.constant(0);
	std::ostringstream cap2;
	old = std::cout.rdbuf(cap2.rdbuf());
	HDKB200::solveGeometricConjugateGradient(pF, rhs, A2, M2, dot2, nrm2, axpy2, addS2, 1e-6, 1000);
	std::cout.rdbuf(old);
	int itF = -1;
	parseHistory(cap2.str(), &itF);
	EXPECT(std::abs(itF - itR) <= 1, "functor-form iterations %d vs %d", itF, itR);
	EXPECT(relDiff(pF, pR) <= 1e-5, "functor-form pressure %.3e", relDiff(pF, pR));
	std::printf("functor form: iterations %d, pressure %.2e\n", itF, relDiff(pF, pR));
    }

    // ---- the production configuration: useGaussSeidel = true (GFS.cpp:463-466, Test.cpp:805-808) -------------------------------
    {
	UT_VoxelArray<double> rhs;
	randomActive(rhs, labR, 9);
	Ref::scaleVector<double>(rhs, dx * dx, labR);
	std::ostringstream quiet;
	std::streambuf *old = std::cout.rdbuf(quiet.rdbuf());
	HDK::GeometricMultigridPoissonSolver mgR(labR, wR, mgLevels, true);
	std::cout.rdbuf(old);
	HDKB200::GeometricMultigridPoissonSolver mgN(labN, wN, mgLevels, true);
	UT_VoxelArray<double> zR = rhs, zN = rhs;
	zR.constant(0); zN.constant(0);
	mgR.applyVCycle(zR, rhs);
	mgN.applyVCycle(zN, rhs);
	EXPECT(relDiff(zN, zR) <= 1e-11, "applyVCycle (Gauss-Seidel) %.3e", relDiff(zN, zR));
	auto A = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &s) { Ref::applyPoissonMatrix<double>(d, s, labR, &wR); };
	auto M = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &s) { mgR.applyVCycle(d, s); };
	auto dot = [&](const UT_VoxelArray<double> &a, const UT_VoxelArray<double> &c) { return Ref::dotProduct<double>(a, c, labR); };
	auto nrm = [&](const UT_VoxelArray<double> &a) { return Ref::squaredL2Norm<double>(a, labR); };
	auto axpy = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &s, double sc) { Ref::addToVector<double>(d, s, sc, labR); };
	auto addS = [&](UT_VoxelArray<double> &d, const UT_VoxelArray<double> &u, const UT_VoxelArray<double> &s, double sc) {
	    Ref::addVectors<double>(d, u, s, sc, labR);
	};
	UT_VoxelArray<double> pR = rhs, pN = rhs;
	pR.constant(0); pN.constant(0);
	std::ostringstream cap;
	old = std::cout.rdbuf(cap.rdbuf());
	HDK::solveGeometricConjugateGradient(pR, rhs, A, M, dot, nrm, axpy, addS, 1e-6, 1000);
	std::cout.rdbuf(old);
	int itR = -1;
	const std::vector<double> hR = parseHistory(cap.str(), &itR);
	std::vector<double> hN;
	const int itN = HDKB200::solveGeometricConjugateGradient(mgN, pN, rhs, 1e-6, 1000, true, &hN);
	EXPECT(std::abs(itR - itN) <= 1, "Gauss-Seidel PCG iterations %d vs %d", itR, itN);
	const size_t m = std::min(hR.size(), hN.size());
	double worst = 0;
	for (size_t i = 0; i < m; ++i) worst = std::max(worst, std::fabs(hR[i] - hN[i]) / hR[i]);
	EXPECT(m > 0 && worst <= 1e-5 + 6e-6, "Gauss-Seidel residual history deviates by %.3e", worst);
	EXPECT(relDiff(pN, pR) <= 1e-5, "Gauss-Seidel pressure %.3e", relDiff(pN, pR));
	std::printf("Gauss-Seidel solver: V-cycle %.2e, PCG iterations %d vs %d, history dev %.2e, pressure %.2e\n", relDiff(zN, zR), itR, itN, worst,
		    relDiff(pN, pR));
    }
    std::printf(g_fail ? "FAILED: %d check(s)\n" : "facade parity ok (%d failures)\n", g_fail);
    return g_fail ? 1 : 0;
}
