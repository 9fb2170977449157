// gmg_band_tiles.cuh -- a whole group of band sweeps in ONE launch with NO barrier between the sweeps.
//
// The band smoother (boundaryJacobiPoissonSmoother, Ops.h:524-619, called myBoundarySmootherIterations = 3 times before and
// after every interior sweep, MG.cpp:445-513) is the V-cycle's launch-latency problem: twelve dependent sweeps per level
// over a list that lives in L2, 4-7 us each whatever the list size.  Replacing the kernel boundaries by grid barriers did
// not pay (k_band_group, k_band_resident: a barrier over ~300 CTAs costs what a programmatic kernel boundary costs).
// Here the dependency is cut instead: the band list is partitioned into spatially compact TILES (Morton order of 4^3 bricks,
// equal cuts), and a tile carries the two rings of band cells around it -- ring 1 = band neighbours of its own cells,
// ring 2 = band neighbours of ring 1.  One CTA per tile then runs
//     sweep 1 on own + ring 1 + ring 2   (from the grid),
//     sweep 2 on own + ring 1            (from shared memory),
//     sweep 3 on own                     (from shared memory),
// which is what three global sweeps compute for the own cells: redundant arithmetic on the rings (~+50 % on sweep 1, +25 %
// on sweep 2 for ~1400-cell tiles) buys a group without any inter-CTA dependency.  Per cell the arithmetic and its order are
// those of bandBody (gmg_kernels.cuh), so the result is bitwise the one of the sweep-per-launch kernels.
// The one remaining hazard is the write-back: sweep 1 of a neighbouring tile reads this tile's own cells from the grid.
// Every CTA therefore ARRIVES on a counter after its sweep 1 and checks the counter before its write-back, two sweeps later
// -- by then everybody has normally arrived and the check is one L2 read.  This needs the tiles co-resident (the host asks
// the occupancy calculator and keeps the sweep-per-launch kernels otherwise).  The first group of a down-stroke (grid known
// to be zero: sweep 1 reads no solution values at all) has no hazard, no counter and no residency requirement.
#pragma once

#include "gmg_kernels.cuh"

namespace gmg
{
constexpr int BT_THREADS = 512;
constexpr unsigned short TREF_SKIP = 0xFFFFu, TREF_FROZEN = 0xFFFEu;
constexpr int BT_MAX_LOC = 0xFFF0;

struct BandTileArgs
{
    double *x;                      // grid
    const double *b;                // grid rhs
    const int4 *tiles;              // per tile: {first local cell, own cells, own + ring 1, own + ring 1 + ring 2}
    const int32_t *locGi;           // grid index of every local cell
    const int32_t *locJ;            // its position in the band list (coefficient records)
    const unsigned short *locCode;  // 2 bits per neighbour: 0 = no coupling, 1 = coefficient 1, 2 = fractional coefficient (bcoef)
    const double *locDiag;
    const unsigned short *locRef;   // tile t, neighbour n, cell c < own + ring 1: at 6 * first + n * (own + ring 1) + c
    const double *bcoef;
    GroupBarrier *bar;
    int nBoundary;
    int pitch;
    int64_t plane;
    int maxLoc, maxCalc;
};

__device__ __forceinline__ double tileCoef(const BandTileArgs &a, int q, int n)
{
    return a.bcoef[int64_t(n) * a.nBoundary + a.locJ[q]];
}

template <bool ZEROGRID>
__global__ void __launch_bounds__(BT_THREADS, 2) k_band_tile(const BandTileArgs a)
{
    pdlLaunch();
    extern __shared__ __align__(16) unsigned char tileSmem[];
    double *v1 = reinterpret_cast<double *>(tileSmem);                       // [maxLoc]
    double *v2 = v1 + a.maxLoc;                                               // [maxCalc]
    double *rhs = v2 + a.maxCalc;                                             // [maxCalc]
    int *gi = reinterpret_cast<int *>(rhs + a.maxCalc);                       // [maxLoc]
    unsigned short *code = reinterpret_cast<unsigned short *>(gi + a.maxLoc);  // [maxLoc]
    unsigned short *ref = code + a.maxLoc;                                    // [6][maxCalc]
    const int4 t = a.tiles[blockIdx.x];
    const int base = t.x, nOwn = t.y, nCalc = t.z, nLoc = t.w;
    const int tid = threadIdx.x;
    // static tables, before the predecessor has finished
    for (int c = tid; c < nLoc; c += BT_THREADS)
    {
	gi[c] = a.locGi[base + c];
	code[c] = a.locCode[base + c];
    }
    {
	const unsigned short *g = a.locRef + int64_t(6) * base;
#pragma unroll
	for (int n = 0; n < 6; ++n)
	    for (int c = tid; c < nCalc; c += BT_THREADS) ref[n * a.maxCalc + c] = g[n * nCalc + c];
    }
    pdlWait();
    const int64_t stride[6] = {-1, 1, -int64_t(a.pitch), int64_t(a.pitch), -a.plane, a.plane};
    // sweep 1: grid -> v1, every local cell
    for (int c = tid; c < nLoc; c += BT_THREADS)
    {
	const int64_t i = gi[c];
	const double rh = a.b[i];
	if (c < nCalc) rhs[c] = rh;
	const double diag = a.locDiag[base + c];
	double centre = 0.0, lap = 0.0;
	if (!ZEROGRID)
	{
	    centre = a.x[i];
	    const unsigned cd = code[c];
#pragma unroll
	    for (int n = 0; n < 6; ++n)
	    {
		const unsigned cc = (cd >> (2 * n)) & 3u;
		if (cc == 0u) continue;
		const double u = a.x[i + stride[n]];
		if (cc == 1u) lap -= u;
		else lap -= tileCoef(a, base + c, n) * u;
	    }
	    lap += diag * centre;
	}
	double r = rh - lap;
	r /= diag;
	v1[c] = centre + (2.0 / 3.0) * r;
    }
    __syncthreads();
    unsigned gen0 = 0;
    if (!ZEROGRID && tid == 0)
    {
	// this CTA has read everything it reads of other tiles' cells: arrive (the last one opens the write-backs)
	gen0 = *reinterpret_cast<volatile unsigned *>(&a.bar->generation);
	if (atomicAdd(&a.bar->count, 1u) == gridDim.x - 1)
	{
	    a.bar->count = 0u;
	    __threadfence();
	    atomicAdd(&a.bar->generation, 1u);
	}
    }
    // sweep 2: v1 -> v2, own + ring 1
    for (int c = tid; c < nCalc; c += BT_THREADS)
    {
	const double diag = a.locDiag[base + c];
	const double centre = v1[c];
	const unsigned cd = code[c];
	double lap = 0.0;
#pragma unroll
	for (int n = 0; n < 6; ++n)
	{
	    const unsigned cc = (cd >> (2 * n)) & 3u;
	    if (cc == 0u) continue;
	    const unsigned short lr = ref[n * a.maxCalc + c];
	    double u;
	    if (lr == TREF_FROZEN)
	    {
		if (ZEROGRID) continue;  // a frozen neighbour holds 0
		u = a.x[int64_t(gi[c]) + stride[n]];
	    }
	    else u = v1[lr];
	    if (cc == 1u) lap -= u;
	    else lap -= tileCoef(a, base + c, n) * u;
	}
	lap += diag * centre;
	double r = rhs[c] - lap;
	r /= diag;
	v2[c] = centre + (2.0 / 3.0) * r;
    }
    __syncthreads();
    // sweep 3: v2 -> (v1, then the grid), own cells
    for (int c = tid; c < nOwn; c += BT_THREADS)
    {
	const double diag = a.locDiag[base + c];
	const double centre = v2[c];
	const unsigned cd = code[c];
	double lap = 0.0;
#pragma unroll
	for (int n = 0; n < 6; ++n)
	{
	    const unsigned cc = (cd >> (2 * n)) & 3u;
	    if (cc == 0u) continue;
	    const unsigned short lr = ref[n * a.maxCalc + c];
	    double u;
	    if (lr == TREF_FROZEN)
	    {
		if (ZEROGRID) continue;
		u = a.x[int64_t(gi[c]) + stride[n]];
	    }
	    else u = v2[lr];
	    if (cc == 1u) lap -= u;
	    else lap -= tileCoef(a, base + c, n) * u;
	}
	lap += diag * centre;
	double r = rhs[c] - lap;
	r /= diag;
	const double v = centre + (2.0 / 3.0) * r;
	if (ZEROGRID) a.x[gi[c]] = v;
	else v1[c] = v;  // same thread reads it back below
    }
    if (!ZEROGRID)
    {
	if (tid == 0)
	{
	    unsigned long long t0 = 0;
	    unsigned spins = 0;
	    while (*reinterpret_cast<volatile unsigned *>(&a.bar->generation) == gen0)
	    {
		if ((++spins & 1023u) == 0)
		{
		    unsigned long long now;
		    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
		    if (t0 == 0) t0 = now;
		    else if (now - t0 > 2000000000ull) { atomicExch(&a.bar->error, 1); break; }
		}
	    }
	}
	__syncthreads();
	for (int c = tid; c < nOwn; c += BT_THREADS) a.x[gi[c]] = v1[c];
    }
}

// ---- tile construction (solver creation) ----
__device__ __forceinline__ unsigned spread10(unsigned v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x30000ffu;
    v = (v | (v << 8)) & 0x300f00fu;
    v = (v | (v << 4)) & 0x30c30c3u;
    v = (v | (v << 2)) & 0x9249249u;
    return v;
}
// sort key of a band cell: Morton code of its 4 x 4 x 4 brick
__global__ void __launch_bounds__(BLOCK) k_tile_keys(unsigned *key, int32_t *val, const int32_t *bandIdx, int nBand, int pitch, int64_t plane)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nBand) return;
    const int64_t i = bandIdx[k];
    const int z = int(i / plane);
    const int rem = int(i - int64_t(z) * plane);
    const int y = rem / pitch, x = rem - y * pitch;
    key[k] = spread10(unsigned(x) >> 2) | (spread10(unsigned(y) >> 2) << 1) | (spread10(unsigned(z) >> 2) << 2);
    val[k] = k;
}
__global__ void __launch_bounds__(BLOCK) k_tile_assign(int32_t *tileOf, int32_t *sortedPos, const int32_t *order, int nBand, int own)
{
    const int p = blockIdx.x * BLOCK + threadIdx.x;
    if (p >= nBand) return;
    const int k = order[p];
    tileOf[k] = p / own;
    sortedPos[k] = p;
}
constexpr unsigned long long TKEY_NONE = ~0ull;
// ring-1 candidates: (tile << 33 | band position << 1 | 0) for every band neighbour that belongs to another tile
__global__ void __launch_bounds__(BLOCK) k_tile_ring1(unsigned long long *keys, const int32_t *bandRef, const int32_t *tileOf, int nBand)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nBand) return;
    const unsigned long long t = unsigned(tileOf[k]);
#pragma unroll
    for (int n = 0; n < 6; ++n)
    {
	const int j = bandRef[int64_t(n) * nBand + k];
	keys[int64_t(n) * nBand + k] = (j >= 0 && unsigned(tileOf[j]) != t) ? ((t << 33) | (unsigned long long)(unsigned(j)) << 1) : TKEY_NONE;
    }
}
// ring-2 candidates from the ring-1 list (flag bit 1), behind a copy of the ring-1 list itself (flag bit 0: wins the tie)
__global__ void __launch_bounds__(BLOCK) k_tile_ring2(unsigned long long *keys, const unsigned long long *ring1, int n1, const int32_t *bandRef,
						     const int32_t *tileOf, int nBand)
{
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n1) return;
    const unsigned long long e = ring1[i];
    const unsigned long long t = e >> 33;
    const int j = int((e >> 1) & 0xffffffffull);
    keys[i] = e;
#pragma unroll
    for (int n = 0; n < 6; ++n)
    {
	const int m = bandRef[int64_t(n) * nBand + j];
	keys[int64_t(n + 1) * n1 + i] = (m >= 0 && unsigned(tileOf[m]) != t) ? ((t << 33) | ((unsigned long long)(unsigned(m)) << 1) | 1ull) : TKEY_NONE;
    }
}
// first entry of every run of equal (key >> shift), the empty key excluded
__global__ void __launch_bounds__(BLOCK) k_tile_heads(uint8_t *head, const unsigned long long *keys, int64_t n, int shift)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= n) return;
    const unsigned long long e = keys[i];
    head[i] = uint8_t(e != TKEY_NONE && (i == 0 || (keys[i - 1] >> shift) != (e >> shift)));
}
// (tile, position, ring) -> (tile, ring, position): the halo of a tile sorted ring 1 first; and the ring sizes per tile
__global__ void __launch_bounds__(BLOCK) k_tile_rekey(unsigned long long *keys, int *count, int n)
{
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const unsigned long long e = keys[i];
    const unsigned long long t = e >> 33, ring = e & 1ull, j = (e >> 1) & 0xffffffffull;
    keys[i] = (t << 33) | (ring << 32) | j;
    atomicAdd(count + 2 * int(t) + int(ring), 1);
}
__device__ __forceinline__ int tileFind(const unsigned long long *halo, int lo, int hi, unsigned long long key)
{
    int a = lo, b = hi;
    while (a < b)
    {
	const int m = (a + b) >> 1;
	if (halo[m] < key) a = m + 1;
	else b = m;
    }
    return (a < hi && halo[a] == key) ? a : -1;
}
struct TileFillArgs
{
    const int4 *tiles;
    const int32_t *haloStart;
    const unsigned long long *halo;
    const int32_t *order, *tileOf, *sortedPos, *bandIdx, *bandRef;
    const double *bcoef;
    const unsigned short *wcode;
    int32_t *locGi, *locJ;
    unsigned short *locCode, *locRef;
    double *locDiag;
    int *error;
    int nTiles, own, nBand, nBoundary, hasWeights, totalLoc;
};
__global__ void __launch_bounds__(BLOCK) k_tile_fill(const TileFillArgs a)
{
    const int q = blockIdx.x * BLOCK + threadIdx.x;
    if (q >= a.totalLoc) return;
    int lo = 0, hi = a.nTiles - 1;  // last tile whose first local cell is <= q
    while (lo < hi)
    {
	const int m = (lo + hi + 1) >> 1;
	if (a.tiles[m].x <= q) lo = m;
	else hi = m - 1;
    }
    const int t = lo;
    const int4 tl = a.tiles[t];
    const int c = q - tl.x;
    const int hs = a.haloStart[t];
    const int j = c < tl.y ? a.order[int64_t(t) * a.own + c] : int(a.halo[hs + c - tl.y] & 0xffffffffull);
    a.locJ[q] = j;
    a.locGi[q] = a.bandIdx[j];
    const bool weighted = a.hasWeights && j < a.nBoundary;
    a.locDiag[q] = j < a.nBoundary ? a.bcoef[int64_t(6) * a.nBoundary + j] : 6.0;
    const unsigned wc = weighted ? a.wcode[j] : 0x555u;
    unsigned cd = 0;
    const int nCalc = tl.z, r1 = tl.z - tl.y;
#pragma unroll
    for (int n = 0; n < 6; ++n)
    {
	const int r = a.bandRef[int64_t(n) * a.nBand + j];
	unsigned short lr = TREF_SKIP;
	if (r != BAND_SKIP)
	{
	    cd |= ((wc >> (2 * n)) & 3u) << (2 * n);
	    if (r < 0) lr = TREF_FROZEN;
	    else if (c < nCalc)
	    {
		int at = -1;
		if (a.tileOf[r] == t) at = a.sortedPos[r] - t * a.own;
		else
		{
		    const unsigned long long k1 = ((unsigned long long)(unsigned(t)) << 33) | unsigned(r);
		    int f = tileFind(a.halo, hs, hs + r1, k1);
		    if (f < 0) f = tileFind(a.halo, hs + r1, hs + (tl.w - tl.y), k1 | (1ull << 32));
		    if (f >= 0) at = tl.y + (f - hs);
		}
		if (at < 0) { atomicExch(a.error, 1); at = 0; }
		lr = (unsigned short)(at);
	    }
	}
	if (c < nCalc) a.locRef[int64_t(6) * tl.x + int64_t(n) * nCalc + c] = lr;
    }
    a.locCode[q] = (unsigned short)(cd);
}
}  // namespace gmg
