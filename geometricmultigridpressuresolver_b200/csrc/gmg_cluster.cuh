// gmg_cluster.cuh -- the coarse sub-V-cycle in ONE thread-block cluster with its vectors in distributed shared memory.
//
// Measured on B200 (profiles/): a V-cycle step on a level of a few thousand to ~100k cells costs 4-5 us as a kernel of its own
// -- launch, a chain of dependent L2 loads, drain -- whatever its size, and the three levels below level 1 of the 256^3
// problem spend 17 such steps each: 40 % of the V-cycle for 2 % of the cells.  Here levels [first, last] -- down-stroke,
// direct solve, up-stroke -- run inside one kernel of ONE cluster (16 CTAs x 1024 threads, or 8 where 16 is refused):
//   * every level's cells are a compact list in storage order, cut into equal blocks, one per CTA; the level's vectors
//     (two ping-pong solution arrays and the right-hand side) live in the OWNING CTA's shared memory;
//   * a neighbour / restriction tap / prolongation corner is a packed (owner CTA, index) reference; a value owned by this
//     CTA is a plain shared-memory load, a value owned by another CTA a distributed-shared-memory load over the SM-to-SM
//     network (cluster.map_shared_rank) -- measured first version (profiles/r02_cluster_cycle.md): sending the LOCAL loads
//     through the cluster window too made level 2 DSMEM-bandwidth-bound (~20 B/cycle/SM), 6 us per sweep;
//   * the neighbour tables of the smoother (16-bit codes: same / previous / next CTA + index, else a "far" escape into the
//     32-bit table), diagonals and flags are copied into shared memory ONCE, in the prologue before griddepcontrol.wait, so
//     a smoothing step touches no global memory at all; the restriction / prolongation tables, used once per V-cycle, are
//     read from global memory (coalesced, L2-resident);
//   * a step ends on the hardware cluster barrier (barrier.cluster arrive.release / wait.acquire) instead of a kernel boundary.
// Per cell the arithmetic and its order are those of k_stencil / k_band / k_restrict / k_prolong / k_coarse_solve (and of
// k_compact_cycle, the one-CTA predecessor of this kernel, kept as the fallback), so the result is bitwise identical.
// Reference semantics: applyVCycle, MG.cpp:557-784 (the levels below the finest), smoothers Ops.h:262-367 / :524-619,
// transfer operators Ops.h:734-972, direct solve MG.cpp:669-692.
#pragma once

#include "gmg_kernels.cuh"

namespace gmg
{
constexpr int CLUSTER_THREADS = 512;
constexpr int CLUSTER_MAX_LEVELS = 10;
constexpr unsigned CLUSTER_NONE = 0xffffffffu;
constexpr int CLUSTER_OWNER_SHIFT = 20;  // packed reference = (owner CTA << 20) | index in the owner's block

// 16-bit neighbour code: bits 15..13 = where, bits 12..0 = index in the owner's block (a block holds at most 8192 cells)
constexpr unsigned short CODE_SAME = 0u << 13, CODE_PREV = 1u << 13, CODE_NEXT = 2u << 13, CODE_NONE = 3u << 13, CODE_FAR = 7u << 13;
// Largest block the cycle takes: two cells per thread.  Measured (profiles/r02_cluster_cycle.md): the 67k-cell level 2 of the
// 256^3 problem, 4.2k cells per CTA, costs 6 us per sweep on 16 SMs (fp64 divisions and shared-memory traffic of 4k cells on
// one SM) against 4.4 us as a kernel of its own over 148 SMs -- the cluster pays only below ~30k cells.
constexpr int CLUSTER_MAX_PER = 1024;
constexpr int CLUSTER_SOLO_MAX = 2048;  // a level of at most this many cells lives in CTA 0 alone (4 cells per thread)

struct ClusterLevel
{
    int n;                 // active cells of the level
    int per;               // cells per CTA block (the last blocks may be short or empty)
    int off;               // offset (doubles) of this level's [xa | xb | b] arrays, `per` each, in every CTA's shared memory
    int tabOff;            // offset (bytes) of this level's shared-memory tables: u16 nbr[6][per], u8 diag[per], u8 flags[per]
    const unsigned short *nbr16;  // [6][n] 16-bit codes of the six neighbours (staged into shared memory block by block)
    const unsigned *nbr;   // [6][n]  packed reference of the -x,+x,-y,+y,-z,+z neighbour, CLUSTER_NONE = not active
    const unsigned *rst;   // [64][n] packed references (next FINER level) of the 4x4x4 restriction taps (levels below the top)
    const unsigned *pro;   // [8][n]  packed references (next COARSER level) of the 2x2x2 prolongation corners
    const uint8_t *diag;   // [n] 6 for INTERIOR; number of non-EXTERIOR neighbours for BOUNDARY (Ops.h:237-248, weight 1)
    const uint8_t *flags;  // [n] bit0 = boundary band, bits 1..3 = parity of the x, y, z storage index
};

struct ClusterArgs
{
    ClusterLevel lv[CLUSTER_MAX_LEVELS];  // lv[0] is the finest level of the cycle
    int nLevels;
    int soloFirst;            // levels [soloFirst, nLevels) live entirely in CTA 0 and step on __syncthreads (nLevels = none)
    int soloCur[CLUSTER_MAX_LEVELS];  // which ping-pong array holds a solo level's solution after its sub-cycle (known from the sweep count)
    int sweeps;
    int scratchOff;           // offset (doubles) of the direct solve's gathered right-hand side in CTA 0
    const int32_t *cellTop;   // [lv[0].n] storage index of lv[0]'s cells in its grid
    const double *bTop;       // rhs grid of lv[0] (written by the restriction kernel of the level above)
    double *xTop;             // solution grid of lv[0] (read by the prolongation kernel of the level above)
    const unsigned *solveRef; // [nSolve] packed reference (coarsest level) of the direct solve's k-th unknown (MG.cpp:296-323 numbering)
    int nSolve;
    const double *inv;        // [nSolve][nSolve]
};

// the value behind a packed reference: `base` is the array's address in THIS CTA's shared memory; every CTA lays its
// shared memory out identically, so the same offset is valid in the owner's
// (a value this CTA owns is a plain shared-memory load: the cluster window moves ~20 bytes per cycle and SM, shared memory 128)
__device__ __forceinline__ double clusterLoad(cg::cluster_group &cl, double *base, unsigned ref, int rank)
{
    const unsigned owner = ref >> CLUSTER_OWNER_SHIFT, idx = ref & ((1u << CLUSTER_OWNER_SHIFT) - 1u);
    if (int(owner) == rank) return base[idx];
    return *cl.map_shared_rank(base + idx, owner);
}

// the level's shared-memory tables in this CTA
struct ClusterTab
{
    const unsigned short *nbr;  // [6][per]
    const uint8_t *diag, *flags;
};
__device__ __forceinline__ ClusterTab clusterTab(const ClusterLevel &L, unsigned char *smBytes)
{
    ClusterTab t;
    t.nbr = reinterpret_cast<const unsigned short *>(smBytes + L.tabOff);
    t.diag = smBytes + L.tabOff + size_t(12) * L.per;
    t.flags = t.diag + L.per;
    return t;
}

// A x at this CTA's j-th cell of the level (compact index k), in the operation order of computeLaplacian (Ops.h:177-260):
// neighbours by (axis, direction), centre last.  x: the solution array in THIS CTA's shared memory.
__device__ __forceinline__ double clusterLap(cg::cluster_group &cl, const ClusterLevel &L, const ClusterTab &T, double *__restrict__ x, int rank, int k, int j,
					     double diag)
{
    unsigned short code[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) code[d] = T.nbr[d * L.per + j];
    double lap = 0.0;
#pragma unroll
    for (int d = 0; d < 6; ++d)
    {
	const unsigned where = code[d] & 0xe000u, idx = code[d] & 0x1fffu;
	if (where == CODE_NONE) continue;
	double u;
	if (where == CODE_SAME) u = x[idx];
	else if (where == CODE_PREV) u = *cl.map_shared_rank(x + idx, rank - 1);
	else if (where == CODE_NEXT) u = *cl.map_shared_rank(x + idx, rank + 1);
	else u = clusterLoad(cl, x, __ldg(L.nbr + size_t(d) * L.n + k), rank);  // far: more than one block away (not at the sizes this cycle takes)
	lap -= u;
    }
    lap += diag * x[j];
    return lap;
}

// one damped-Jacobi sweep xin -> xout over the band cells (bandOnly; everything else is copied) or over all cells
// (Ops.h:262-367, :524-619); the caller ends the step (cluster barrier, or __syncthreads inside the solo sub-cycle)
__device__ __forceinline__ void clusterSweep(cg::cluster_group &cl, const ClusterLevel &L, const ClusterTab &T, double *__restrict__ xin, double *__restrict__ xout,
					     const double *__restrict__ b, int rank, int first, int count, bool bandOnly)
{
#pragma unroll 2
    for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS)
    {
	if (bandOnly && !(T.flags[j] & 1)) { xout[j] = xin[j]; continue; }
	const double diag = double(T.diag[j]);
	const double lap = clusterLap(cl, L, T, xin, rank, first + j, j, diag);
	double r = b[j] - lap;
	r /= diag;
	xout[j] = xin[j] + (2.0 / 3.0) * r;
    }
}

// Levels [soloFirst, nLevels) are small enough (<= CLUSTER_SOLO_MAX cells) to live entirely in CTA 0: there a step ends on
// __syncthreads (0.08 us) instead of the cluster barrier (0.5 us with 16 x 512 threads, scripts/cluster_probe.cu), every access
// is a plain shared-memory access, and the other CTAs wait at ONE cluster barrier for the whole sub-cycle.
__global__ void __launch_bounds__(CLUSTER_THREADS, 1) k_cluster_cycle(const ClusterArgs c)
{
    pdlLaunch();
    extern __shared__ double sm[];
    unsigned char *smBytes = reinterpret_cast<unsigned char *>(sm);
    cg::cluster_group cl = cg::this_cluster();
    const int rank = int(cl.block_rank());
    const int nl = c.nLevels;
    // this CTA's block of level q: cells [first, first + count)
    auto blockOf = [&](int q, int &first, int &count) {
	const ClusterLevel &L = c.lv[q];
	first = rank * L.per;
	count = max(0, min(L.per, L.n - first));
    };
    // solution arrays ping-pong: cur[q] tells which of xa / xb holds the level's current solution
    int cur[CLUSTER_MAX_LEVELS];
    auto X = [&](int q, int which) { return sm + c.lv[q].off + which * c.lv[q].per; };
    auto B = [&](int q) { return sm + c.lv[q].off + 2 * c.lv[q].per; };
    // end of a step: the cluster barrier, or the CTA barrier inside the solo sub-cycle
    auto stepSync = [&](bool solo) {
	if (solo) __syncthreads();
	else cl.sync();
    };
    // band sweeps, one interior sweep, band sweeps (MG.cpp:445-513 and its per-level copies)
    auto smooth = [&](int q, bool solo) {
	const ClusterLevel &L = c.lv[q];
	int first, count;
	blockOf(q, first, count);
	const ClusterTab T = clusterTab(L, smBytes);
	for (int s = 0; s < 2 * c.sweeps + 1; ++s)
	{
	    clusterSweep(cl, L, T, X(q, cur[q]), X(q, cur[q] ^ 1), B(q), rank, first, count, s != c.sweeps);
	    stepSync(solo);
	    cur[q] ^= 1;
	}
    };
    // down-stroke of level q (MG.cpp:557-667): x = 0, smooth, residual, restrict into level q + 1
    auto downLevel = [&](int q, bool solo) {
	const ClusterLevel &L = c.lv[q];
	const ClusterLevel &C = c.lv[q + 1];
	int first, count;
	blockOf(q, first, count);
	cur[q] = 0;
	{
	    double *x = X(q, 0);
	    for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS) x[j] = 0.0;
	}
	stepSync(solo);
	smooth(q, solo);
	{
	    // residual into the other solution array (the current one is kept for the up-stroke)
	    double *x = X(q, cur[q]), *t = X(q, cur[q] ^ 1);
	    const double *b = B(q);
	    const ClusterTab T = clusterTab(L, smBytes);
#pragma unroll 2
	    for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS)
	    {
		const double lap = clusterLap(cl, L, T, x, rank, first + j, j, double(T.diag[j]));
		t[j] = b[j] + (-1.0) * lap;  // Ops.h:731
	    }
	}
	stepSync(solo);
	{
	    int cfirst, ccount;
	    blockOf(q + 1, cfirst, ccount);
	    double *t = X(q, cur[q] ^ 1);
	    double *bc = B(q + 1);
	    for (int j = threadIdx.x; j < ccount; j += CLUSTER_THREADS)
	    {
		const int k = cfirst + j;
		const double rw[4] = {1. / 8., 3. / 8., 3. / 8., 1. / 8.};
		double v = 0.0;
		// one z-slice of taps per round: 16 independent table loads, then 16 independent value loads
#pragma unroll
		for (int z = 0; z < 4; ++z)
		{
		    unsigned ref[16];
#pragma unroll
		    for (int t16 = 0; t16 < 16; ++t16) ref[t16] = __ldg(C.rst + size_t(z * 16 + t16) * C.n + k);
		    double val[16];
#pragma unroll
		    for (int t16 = 0; t16 < 16; ++t16) val[t16] = ref[t16] != CLUSTER_NONE ? clusterLoad(cl, t, ref[t16], rank) : 0.0;
#pragma unroll
		    for (int y = 0; y < 4; ++y)
#pragma unroll
			for (int xx = 0; xx < 4; ++xx)
			    if (ref[y * 4 + xx] != CLUSTER_NONE) v += rw[xx] * rw[y] * rw[z] * val[y * 4 + xx];  // an inactive tap adds +0.0, which never changes v
		}
		bc[j] = v;
	    }
	}
	stepSync(solo);
    };
    // direct solve on the coarsest level (MG.cpp:669-692): x = A^-1 b, one warp per row as in k_coarse_solve, in CTA 0
    auto solveCoarsest = [&](bool solo) {
	const int q = nl - 1;
	cur[q] = 0;
	double *x = X(q, 0);
	if (rank == 0)
	{
	    double *sb = sm + c.scratchOff;
	    const int n = c.nSolve;
	    for (int i = threadIdx.x; i < n; i += CLUSTER_THREADS) sb[i] = clusterLoad(cl, B(q), __ldg(c.solveRef + i), rank);
	    __syncthreads();
	    const int lane = threadIdx.x & 31;
	    for (int row = threadIdx.x >> 5; row < n; row += CLUSTER_THREADS / 32)
	    {
		const double *r = c.inv + int64_t(row) * n;
		double acc = 0.0;
		for (int j = lane; j < n; j += 32) acc += r[j] * sb[j];
		acc = warpSum(acc);
		if (lane == 0)
		{
		    const unsigned ref = __ldg(c.solveRef + row);
		    const unsigned owner = ref >> CLUSTER_OWNER_SHIFT, idx = ref & ((1u << CLUSTER_OWNER_SHIFT) - 1u);
		    if (owner == 0u) x[idx] = acc;
		    else *cl.map_shared_rank(x + idx, owner) = acc;
		}
	    }
	}
	stepSync(solo);
    };
    // up-stroke of level q (MG.cpp:695-784): x += 4 trilerp(x of level q + 1), smooth
    auto upLevel = [&](int q, bool solo) {
	const ClusterLevel &L = c.lv[q];
	int first, count;
	blockOf(q, first, count);
	double *x = X(q, cur[q]);
	double *xc = X(q + 1, cur[q + 1]);
	const ClusterTab T = clusterTab(L, smBytes);
	for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS)
	{
	    const int k = first + j;
	    const int f = T.flags[j];
	    const double wx = (f & 2) ? .25 : .75, wy = (f & 4) ? .25 : .75, wz = (f & 8) ? .25 : .75;
	    unsigned ref[8];
#pragma unroll
	    for (int p = 0; p < 8; ++p) ref[p] = __ldg(L.pro + size_t(p) * L.n + k);
	    double v[8];
#pragma unroll
	    for (int p = 0; p < 8; ++p) v[p] = ref[p] != CLUSTER_NONE ? clusterLoad(cl, xc, ref[p], rank) : 0.0;
	    // corner p = x + 2 y + 4 z; lerp nesting x -> y -> z (Ops.h:841-871)
	    const double e = lerpRef(lerpRef(lerpRef(v[0], v[1], wx), lerpRef(v[2], v[3], wx), wy),
				     lerpRef(lerpRef(v[4], v[5], wx), lerpRef(v[6], v[7], wx), wy), wz);
	    x[j] = x[j] + 4. * e;
	}
	stepSync(solo);
	smooth(q, solo);
    };

    // prologue: this CTA's blocks of the (static) smoother tables go to shared memory while the restriction above the cycle's
    // first level is still running
    for (int q = 0; q < nl; ++q)
    {
	const ClusterLevel &L = c.lv[q];
	int first, count;
	blockOf(q, first, count);
	unsigned short *tn = reinterpret_cast<unsigned short *>(smBytes + L.tabOff);
	uint8_t *td = smBytes + L.tabOff + size_t(12) * L.per, *tf = td + L.per;
	for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS)
	{
#pragma unroll
	    for (int d = 0; d < 6; ++d) tn[d * L.per + j] = __ldg(L.nbr16 + size_t(d) * L.n + first + j);
	    td[j] = __ldg(L.diag + first + j);
	    tf[j] = __ldg(L.flags + first + j);
	}
    }
    pdlWait();
    {
	int first, count;
	blockOf(0, first, count);
	double *b = B(0);
	for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS) b[j] = c.bTop[__ldg(c.cellTop + first + j)];
    }
    const int soloFirst = c.soloFirst;  // levels [soloFirst, nl) live in CTA 0 alone (nl = none)
    // ---- cluster-wide down-stroke; the last restriction lands in CTA 0's block of the first solo level
    for (int q = 0; q < min(soloFirst, nl - 1); ++q) downLevel(q, false);
    if (soloFirst < nl)
    {
	// ---- the solo sub-cycle in CTA 0; everybody else waits at the barrier below
	if (soloFirst == 0) cl.sync();  // (the tables and the top rhs written above are CTA-local, but keep the cluster in step)
	if (rank == 0)
	{
	    for (int q = soloFirst; q + 1 < nl; ++q) downLevel(q, true);
	    solveCoarsest(true);
	    for (int q = nl - 2; q >= soloFirst; --q) upLevel(q, true);
	}
	else
	    for (int q = soloFirst; q < nl; ++q) cur[q] = c.soloCur[q];  // which array holds a solo level's result: fixed by the sweep count
	cl.sync();
    }
    else solveCoarsest(false);
    // ---- cluster-wide up-stroke
    for (int q = min(soloFirst, nl - 1) - 1; q >= 0; --q) upLevel(q, false);
    {
	int first, count;
	blockOf(0, first, count);
	const double *x = X(0, cur[0]);
	for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS) c.xTop[__ldg(c.cellTop + first + j)] = x[j];
    }
}


// ------------------------------------------------------------------------------------------------
// ONE level's smoothing in a cluster (what the measurements above leave standing).  A level of 2k..16k cells costs 4.1 us per
// V-cycle step as a kernel of its own and 17 steps per V-cycle; in a 16-CTA cluster with one cell per thread a step is a
// cluster barrier (~0.45 us) plus a few hundred cycles of shared-memory loads and fp64.  What did NOT pay is moving values
// between a distributed level and its neighbours inside the cluster (one SM's ~20 B/cycle port) -- so here the level's
// right-hand side comes from its grid and its solution / residual go back to their grids, and restriction, the coarser levels
// and prolongation stay the kernels they were.
//   up == 0 (down-stroke, MG.cpp:557-667): x = 0, band sweeps + interior sweep + band sweeps, residual; writes x and r
//   up == 1 (up-stroke,   MG.cpp:695-784): x = grid (the prolongation kernel has added the correction), the same sweeps; writes x
// Same arithmetic, same order as the per-kernel path: bitwise identical (tests/test_gpu_parity.py).
// ------------------------------------------------------------------------------------------------
struct ClusterSmoothArgs
{
    ClusterLevel lv;        // the level (tables as in ClusterArgs)
    int sweeps;
    const int32_t *cell;    // [n] storage index of the compact cells
    const double *b;        // rhs grid
    double *x;              // solution grid
    double *r;              // residual grid (down-stroke)
};

__global__ void __launch_bounds__(CLUSTER_THREADS, 1) k_cluster_smooth(const ClusterSmoothArgs c, int up)
{
    pdlLaunch();
    extern __shared__ double sm[];
    unsigned char *smBytes = reinterpret_cast<unsigned char *>(sm);
    cg::cluster_group cl = cg::this_cluster();
    const int rank = int(cl.block_rank());
    const ClusterLevel &L = c.lv;
    const int first = rank * L.per, count = max(0, min(L.per, L.n - first));
    double *xa = sm + L.off, *xb = xa + L.per, *b = xb + L.per;
    // prologue: this CTA's block of the (static) tables
    {
	unsigned short *tn = reinterpret_cast<unsigned short *>(smBytes + L.tabOff);
	uint8_t *td = smBytes + L.tabOff + size_t(12) * L.per, *tf = td + L.per;
	for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS)
	{
#pragma unroll
	    for (int d = 0; d < 6; ++d) tn[d * L.per + j] = __ldg(L.nbr16 + size_t(d) * L.n + first + j);
	    td[j] = __ldg(L.diag + first + j);
	    tf[j] = __ldg(L.flags + first + j);
	}
    }
    pdlWait();
    for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS)
    {
	const int64_t g = __ldg(c.cell + first + j);
	b[j] = c.b[g];
	xa[j] = up ? c.x[g] : 0.0;
    }
    cl.sync();
    const ClusterTab T = clusterTab(L, smBytes);
    double *cur = xa, *nxt = xb;
    for (int s = 0; s < 2 * c.sweeps + 1; ++s)
    {
	clusterSweep(cl, L, T, cur, nxt, b, rank, first, count, s != c.sweeps);
	cl.sync();
	double *t = cur; cur = nxt; nxt = t;
    }
    for (int j = threadIdx.x; j < count; j += CLUSTER_THREADS)
    {
	const int64_t g = __ldg(c.cell + first + j);
	c.x[g] = cur[j];
	if (!up)
	{
	    const double lap = clusterLap(cl, L, T, cur, rank, first + j, j, double(T.diag[j]));
	    c.r[g] = b[j] + (-1.0) * lap;  // Ops.h:731
	}
    }
    // (no CTA may leave while a neighbour still reads its shared memory for the residual)
    cl.sync();
}

// ------------------------------------------------------------------------------------------------
// table builders (device side: the constructor runs every simulation frame)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_active_flags(uint8_t *flags, const uint8_t *labels, int64_t total)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i < total)
    {
	const int l = labels[i];
	flags[i] = (l == L_INTERIOR || l == L_BOUNDARY) ? 1 : 0;
    }
}

__device__ __forceinline__ unsigned clusterPack(int p, int per)
{
    if (p < 0) return CLUSTER_NONE;
    const int owner = p / per;
    return (unsigned(owner) << CLUSTER_OWNER_SHIFT) | unsigned(p - owner * per);
}

// neighbour references, diagonal and flags of one level (cell: compact -> storage index; pos: storage index -> compact, -1)
__global__ void __launch_bounds__(BLOCK) k_cluster_nbr(unsigned *nbr, uint8_t *diag, uint8_t *flags, const int32_t *cell, const int32_t *pos,
						      const uint8_t *labels, const uint8_t *bandFlags, int n, int per, int pitch, int64_t plane)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= n) return;
    const int64_t i = cell[k];
    const int64_t stride[6] = {-1, 1, -int64_t(pitch), int64_t(pitch), -plane, plane};
    int dg = 0;
#pragma unroll
    for (int d = 0; d < 6; ++d)
    {
	const int64_t j = i + stride[d];
	const int l = labels[j];
	unsigned ref = CLUSTER_NONE;
	if (l == L_INTERIOR || l == L_BOUNDARY) { ref = clusterPack(pos[j], per); ++dg; }
	else if (l == L_DIRICHLET) ++dg;
	nbr[size_t(d) * n + k] = ref;
    }
    diag[k] = uint8_t(dg);  // 6 for an INTERIOR cell (all six neighbours active by construction)
    const int z = int(i / plane);
    const int64_t rem = i - int64_t(z) * plane;
    const int y = int(rem / pitch), x = int(rem - int64_t(y) * pitch);
    flags[k] = uint8_t((bandFlags[i] & 1) | ((x & 1) << 1) | ((y & 1) << 2) | ((z & 1) << 3));
}

// 16-bit neighbour codes from the packed 32-bit references
__global__ void __launch_bounds__(BLOCK) k_cluster_nbr16(unsigned short *nbr16, const unsigned *nbr, int n, int per)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= n) return;
    const int mine = k / per;
#pragma unroll
    for (int d = 0; d < 6; ++d)
    {
	const unsigned ref = nbr[size_t(d) * n + k];
	unsigned short code = CODE_NONE;
	if (ref != CLUSTER_NONE)
	{
	    const int owner = int(ref >> CLUSTER_OWNER_SHIFT);
	    const unsigned idx = ref & ((1u << CLUSTER_OWNER_SHIFT) - 1u);
	    if (owner == mine) code = CODE_SAME | (unsigned short)idx;
	    else if (owner == mine - 1) code = CODE_PREV | (unsigned short)idx;
	    else if (owner == mine + 1) code = CODE_NEXT | (unsigned short)idx;
	    else code = CODE_FAR;
	}
	nbr16[size_t(d) * n + k] = code;
    }
}

// restriction taps of a coarse level in the next finer one (Ops.h:760-834); shift: coarse storage = (fine storage >> 1) + shift
__global__ void __launch_bounds__(BLOCK) k_cluster_rst(unsigned *rst, const int32_t *cellC, const int32_t *posF, int nC, int perF, int pitchC, int64_t planeC,
						      int pitchF, int64_t planeF, int nxF, int nyF, int nzF, int s0, int s1, int s2)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nC) return;
    const int64_t i = cellC[k];
    const int cz = int(i / planeC);
    const int64_t rem = i - int64_t(cz) * planeC;
    const int cy = int(rem / pitchC), cx = int(rem - int64_t(cy) * pitchC);
    const int fx = 2 * (cx - s0) - 1, fy = 2 * (cy - s1) - 1, fz = 2 * (cz - s2) - 1;
    for (int z = 0; z < 4; ++z)
	for (int y = 0; y < 4; ++y)
	    for (int x = 0; x < 4; ++x)
	    {
		const int X = fx + x, Y = fy + y, Z = fz + z;
		unsigned ref = CLUSTER_NONE;
		if (X >= 0 && Y >= 0 && Z >= 0 && X < nxF && Y < nyF && Z < nzF) ref = clusterPack(posF[int64_t(Z) * planeF + int64_t(Y) * pitchF + X], perF);
		rst[size_t((z * 4 + y) * 4 + x) * nC + k] = ref;
	    }
}

// prolongation corners of a fine level in the next coarser one (Ops.h:895-971)
__global__ void __launch_bounds__(BLOCK) k_cluster_pro(unsigned *pro, const int32_t *cellF, const int32_t *posC, int nF, int perC, int pitchF, int64_t planeF,
						      int pitchC, int64_t planeC, int nxC, int nyC, int nzC, int s0, int s1, int s2)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nF) return;
    const int64_t i = cellF[k];
    const int fz = int(i / planeF);
    const int64_t rem = i - int64_t(fz) * planeF;
    const int fy = int(rem / pitchF), fx = int(rem - int64_t(fy) * pitchF);
    const int mx = (fx >> 1) + s0, my = (fy >> 1) + s1, mz = (fz >> 1) + s2;
    const int xs = (fx & 1) ? mx : mx - 1, ys = (fy & 1) ? my : my - 1, zs = (fz & 1) ? mz : mz - 1;
    for (int c = 0; c < 8; ++c)
    {
	const int X = xs + (c & 1), Y = ys + ((c >> 1) & 1), Z = zs + (c >> 2);
	unsigned ref = CLUSTER_NONE;
	if (X >= 0 && Y >= 0 && Z >= 0 && X < nxC && Y < nyC && Z < nzC) ref = clusterPack(posC[int64_t(Z) * planeC + int64_t(Y) * pitchC + X], perC);
	pro[size_t(c) * nF + k] = ref;
    }
}

__global__ void __launch_bounds__(BLOCK) k_cluster_solve_ref(unsigned *ref, const int32_t *coarseIdx, const int32_t *pos, int n, int per)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k < n) ref[k] = clusterPack(pos[coarseIdx[k]], per);
}

} // namespace gmg
