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
// the value type of a kernel: double everywhere on the reference path; float for the V-cycle of the mixed-precision option
// (fp32 V-cycle inside the fp64 CG) -- the stencil / band / transfer bodies are templates on it
template <typename T> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };
template <typename T> __device__ __forceinline__ typename Vec2<T>::type ld2(const T *p) { return *reinterpret_cast<const typename Vec2<T>::type *>(p); }
template <typename T> __device__ __forceinline__ void st2(T *p, typename Vec2<T>::type v) { *reinterpret_cast<typename Vec2<T>::type *>(p) = v; }
template <typename T> __device__ __forceinline__ typename Vec2<T>::type make2(T a, T b) { typename Vec2<T>::type v; v.x = a; v.y = b; return v; }

// Programmatic dependent launch (see launchK in gmg_b200.cu).  A V-cycle at 256^3 is ~70 dependent steps of a few
// microseconds, each a chain of dependent loads (chunk id -> labels / neighbour references -> values), so every hot
// kernel is split in two: a PROLOGUE that lets the next kernel of the chain be scheduled right away (pdlLaunch) and loads
// its own STATIC metadata -- chunk ids, labels, band indices / references / coefficient records, none of which is ever
// written after the constructor -- while the previous kernel is still running, then pdlWait(), after which everything the
// previous kernel wrote is visible and only the value loads remain on the critical path.  Both instructions are no-ops
// for a launch without the PDL attribute.  The "memory" clobbers keep the compiler from moving a load across the wait.
// A/B switch (GMG_PDL_PREFETCH=0): wait right at the top, i.e. no prologue overlap -- the round-1 behaviour
__constant__ int c_pdlWaitFirst = 0;
__device__ __forceinline__ void pdlLaunch()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (c_pdlWaitFirst) asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdlEnter()
{
    pdlLaunch();
    pdlWait();
}

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
    double rho;      // z.r of the previous direction (retired by the update kernel)
    double pAp;      // p.Ap
    double rr;       // |r|^2
    double rhoNew;   // z.r of the current direction
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
// SM_JACOBI_ZERO: the interior sweep of a down-stroke, where x is zero everywhere except on the boundary band (the band
// sweeps before it started from x = 0): the grid is NOT zero-filled first -- whatever it holds off the band is ignored.  A cell
// whose 7-point neighbourhood holds no band cell gets (2/3) b / 6 without reading x at all; near the band every value is
// taken through the band mask.  Bitwise the result of SM_JACOBI on a zero-filled grid.
enum StencilMode { SM_JACOBI = 0, SM_APPLY = 1, SM_RESIDUAL = 2, SM_JACOBI_ZERO = 3 };

template <typename T>
struct StencilArgsT
{
    const uint8_t *labels;
    const uint8_t *flags;   // SM_JACOBI_ZERO: the level's neighbour-mask grid (k_band_nbr_mask): bit0 = the cell is in the boundary band, bits 1..6 = its
			    // -x, +x, -y, +y, -z, +z neighbour is (INTERIOR cells only); 0 = no band cell anywhere in the 7-point neighbourhood
    const T *in;        // x (Jacobi/residual) or source (apply)
    const T *b;         // rhs (Jacobi/residual)
    T *out;             // Jacobi: new x (out of place); apply: A in; residual: b - A in
    const int32_t *chunks;
    int nChunks;
    int chunksPerPlane;
    int pitch;
    int64_t plane;
    int nz;
    int zlo, zhi;       // z-plane clip [zlo, zhi): planes outside are left untouched (deep-halo sharding)
    int dotLo, dotHi;   // planes whose cells enter the fused dot product (the rank's owned planes)
    // boundary records
    int nBoundary;
    const int32_t *bandIdx;
    const double *bcoef;
    const unsigned short *wcode;  // coefficient codes of the BOUNDARY cells (k_band_coef)
    // fused dot(in, A in) for apply
    double *partials;
    unsigned *ticket;
    double *result;
};
typedef StencilArgsT<double> StencilArgs;

template <int MODE, typename T>
__device__ __forceinline__ T stencilFinish(T lap, T centre, T rhs, T diag)
{
    if (MODE == SM_APPLY) return lap;
    if (MODE == SM_RESIDUAL) return rhs + T(-1.0) * lap;  // addVectors(residual, rhs, residual, -1), Ops.h:731
    // SM_JACOBI and SM_JACOBI_ZERO:
    T r = rhs - lap;                                      // Ops.h:357-361
    r /= diag;
    return centre + T(2.0 / 3.0) * r;
}

// BOUNDARY-labelled cell k of the level's record list (Ops.h:208-255): coefficient record in the prologue, values after the wait
template <typename T, int MODE, bool DOT>
__device__ __forceinline__ double stencilBoundary(const StencilArgsT<T> &a, int k)
{
    double acc = 0.0;
    const int64_t i = k < a.nBoundary ? int64_t(a.bandIdx[k]) : 0;
    const int z = int(unsigned(i) / unsigned(a.plane));  // a level's box holds fewer than 2^31 cells: 32-bit division
    const bool live = k < a.nBoundary && z >= a.zlo && z < a.zhi;
    // prologue: the code word, the diagonal and the (rare) fractional coefficients
    const unsigned code = live ? a.wcode[k] : 0u;
    const T diag = live ? T(a.bcoef[int64_t(6) * a.nBoundary + k]) : T(1.0);
    T cn[6];
#pragma unroll
    for (int n = 0; n < 6; ++n)
    {
	const unsigned c = (code >> (2 * n)) & 3u;
	cn[n] = c == 2u ? T(a.bcoef[int64_t(n) * a.nBoundary + k]) : T(c);  // 0, exactly 1, or the fractional coefficient
    }
    pdlWait();
    if (live)
    {
	const int64_t stride[6] = {-1, 1, -int64_t(a.pitch), int64_t(a.pitch), -a.plane, a.plane};
	const T centre = a.in[i];
	T lap = T(0.0);
#pragma unroll
	for (int n = 0; n < 6; ++n)
	    if (cn[n] != T(0.0)) lap -= cn[n] * a.in[i + stride[n]];
	lap += diag * centre;
	const T rhs = (MODE != SM_APPLY) ? a.b[i] : T(0.0);
	a.out[i] = stencilFinish<MODE, T>(lap, centre, rhs, diag);
	if (DOT && z >= a.dotLo && z < a.dotHi) acc += double(centre * lap);
    }
    return acc;
}

// vb = virtual CTA index (blockIdx.x of the stand-alone kernel), tid = thread index inside the BLOCK-wide virtual CTA.
// Returns this thread's part of dot(in, A in) when DOT.  Prologue (before pdlWait): chunk id and the labels of the
// thread's cells (stencilLabels), or the boundary cell's index and coefficient record; then the values (stencilChunk).
struct ChunkLabels
{
    uchar2 lab[CHUNK_Z], flg[CHUNK_Z];
    int64_t inPlane;
    int z0;
};
template <typename T, int MODE>
__device__ __forceinline__ void stencilLabels(const StencilArgsT<T> &a, int vb, int tid, ChunkLabels &cl)
{
    const int c = a.chunks[vb];
    const int zb = c / a.chunksPerPlane;
    cl.inPlane = int64_t(c - zb * a.chunksPerPlane) * CHUNK_CELLS + 2 * tid;
    cl.z0 = zb * CHUNK_Z;
    const int64_t inPlane = cl.inPlane;
    const int z0 = cl.z0;
#pragma unroll
    for (int dz = 0; dz < CHUNK_Z; ++dz)
    {
	const int z = z0 + dz;
	cl.lab[dz] = make_uchar2(L_EXTERIOR, L_EXTERIOR);
	cl.flg[dz] = make_uchar2(0, 0);
	if (inPlane < a.plane && z >= a.zlo && z < a.zhi)
	{
	    cl.lab[dz] = *reinterpret_cast<const uchar2 *>(a.labels + int64_t(z) * a.plane + inPlane);
	    if (MODE == SM_JACOBI_ZERO) cl.flg[dz] = *reinterpret_cast<const uchar2 *>(a.flags + int64_t(z) * a.plane + inPlane);
	}
    }
}
template <typename T, int MODE, bool DOT>
__device__ __forceinline__ double stencilChunk(const StencilArgsT<T> &a, const ChunkLabels &cl)
{
    typedef typename Vec2<T>::type T2;
    double acc = 0.0;
    const int64_t inPlane = cl.inPlane;
    const int z0 = cl.z0;
    const uchar2 *lab = cl.lab, *flg = cl.flg;
    if (MODE == SM_JACOBI_ZERO)
    {
	// One branch-free pass, the four planes' loads in flight together like the other modes.  The mask byte of a cell (prologue)
	// says which of its seven stencil points lie on the band: a point is LOADED only where its bit is set and is zero elsewhere,
	// whatever the grid holds there.  For the common cell -- mask 0, no band cell in the neighbourhood -- every load is predicated
	// off and the arithmetic below runs on zeros: lap = +0.0, x = 0 + (2/3)((b - 0) / 6), a pure stream of b in, x out.  It is the
	// plain sweep's instruction sequence on masked values, hence bitwise its result on a zero-filled grid.
#pragma unroll
	for (int dz = 0; dz < CHUNK_Z; ++dz)
	{
	    const bool a0 = (lab[dz].x == L_INTERIOR), a1 = (lab[dz].y == L_INTERIOR);
	    if (!(a0 | a1)) continue;
	    const int64_t i = int64_t(z0 + dz) * a.plane + inPlane;
	    const T2 rhs = ld2(a.b + i);
	    const unsigned m0 = flg[dz].x, m1 = flg[dz].y, m = m0 | m1;
	    const T2 zero2 = make2<T>(T(0.0), T(0.0));
	    T2 c2 = (m & 1u) ? ld2(a.in + i) : zero2;
	    const T xm = (m0 & 2u) ? a.in[i - 1] : T(0.0), xp = (m1 & 4u) ? a.in[i + 2] : T(0.0);
	    T2 ym = (m & 8u) ? ld2(a.in + i - a.pitch) : zero2, yp = (m & 16u) ? ld2(a.in + i + a.pitch) : zero2;
	    T2 zm = (m & 32u) ? ld2(a.in + i - a.plane) : zero2, zp = (m & 64u) ? ld2(a.in + i + a.plane) : zero2;
	    if (!(m0 & 1u)) c2.x = T(0.0);
	    if (!(m1 & 1u)) c2.y = T(0.0);
	    if (!(m0 & 8u)) ym.x = T(0.0);
	    if (!(m1 & 8u)) ym.y = T(0.0);
	    if (!(m0 & 16u)) yp.x = T(0.0);
	    if (!(m1 & 16u)) yp.y = T(0.0);
	    if (!(m0 & 32u)) zm.x = T(0.0);
	    if (!(m1 & 32u)) zm.y = T(0.0);
	    if (!(m0 & 64u)) zp.x = T(0.0);
	    if (!(m1 & 64u)) zp.y = T(0.0);
	    T lap0 = -xm;
	    lap0 -= c2.y; lap0 -= ym.x; lap0 -= yp.x; lap0 -= zm.x; lap0 -= zp.x;
	    lap0 += T(6.0) * c2.x;
	    T lap1 = -c2.x;
	    lap1 -= xp; lap1 -= ym.y; lap1 -= yp.y; lap1 -= zm.y; lap1 -= zp.y;
	    lap1 += T(6.0) * c2.y;
	    const T o0 = stencilFinish<MODE, T>(lap0, c2.x, rhs.x, T(6.0));
	    const T o1 = stencilFinish<MODE, T>(lap1, c2.y, rhs.y, T(6.0));
	    if (a0 & a1) st2(a.out + i, make2<T>(o0, o1));
	    else if (a0) a.out[i] = o0;
	    else a.out[i + 1] = o1;
	}
	return acc;
    }
#pragma unroll
    for (int dz = 0; dz < CHUNK_Z; ++dz)
    {
	const int z = z0 + dz;
	const uchar2 l = lab[dz];
	const bool a0 = (l.x == L_INTERIOR), a1 = (l.y == L_INTERIOR);
	if (!(a0 | a1)) continue;
	const int64_t i = int64_t(z) * a.plane + inPlane;
	const T2 c2 = ld2(a.in + i);
	const T xm = a.in[i - 1], xp = a.in[i + 2];
	const T2 ym = ld2(a.in + i - a.pitch), yp = ld2(a.in + i + a.pitch);
	const T2 zm = ld2(a.in + i - a.plane), zp = ld2(a.in + i + a.plane);
	T2 rhs = make2<T>(T(0.0), T(0.0));
	if (MODE != SM_APPLY) rhs = ld2(a.b + i);
	T lap0 = -xm;
	lap0 -= c2.y; lap0 -= ym.x; lap0 -= yp.x; lap0 -= zm.x; lap0 -= zp.x;
	lap0 += T(6.0) * c2.x;
	T lap1 = -c2.x;
	lap1 -= xp; lap1 -= ym.y; lap1 -= yp.y; lap1 -= zm.y; lap1 -= zp.y;
	lap1 += T(6.0) * c2.y;
	const T o0 = stencilFinish<MODE, T>(lap0, c2.x, rhs.x, T(6.0));
	const T o1 = stencilFinish<MODE, T>(lap1, c2.y, rhs.y, T(6.0));
	if (a0 & a1) st2(a.out + i, make2<T>(o0, o1));
	else if (a0) a.out[i] = o0;
	else a.out[i + 1] = o1;
	if (DOT && z >= a.dotLo && z < a.dotHi) acc += double((a0 ? c2.x * lap0 : T(0.0)) + (a1 ? c2.y * lap1 : T(0.0)));
    }
    return acc;
}
// The same chunk with the loads of PB planes issued TOGETHER before any of them is used.  stencilChunk above walks its four planes
// one after the other -- loads, arithmetic, store, next plane: the store to `out` may alias the next plane's loads as far as the
// compiler knows, and the division's slow-path call splits the block -- so a CTA is a chain of four dependent round trips, and a
// 256^3 level is 5-6 waves of such CTAs.  Here every value a group of PB planes needs sits in registers first (the centre pairs of
// planes zb - 1 .. zb + PB are loaded once and serve as centre, -z and +z neighbour: 6 plane loads instead of 12 at PB = 4), then
// the group is computed and stored.  Which points are read: all seven, or (SM_JACOBI_ZERO) those the cell's mask byte names.
// Same per-cell arithmetic in the same order, same accumulation order of the fused dot product: bitwise stencilChunk.
template <typename T, int MODE, bool DOT, int PB>
__device__ __forceinline__ double stencilChunkBatched(const StencilArgsT<T> &a, const ChunkLabels &cl)
{
    typedef typename Vec2<T>::type T2;
    static_assert(CHUNK_Z % PB == 0, "plane groups must tile the chunk");
    double acc = 0.0;
    const int64_t inPlane = cl.inPlane;
    const T2 zero2 = make2<T>(T(0.0), T(0.0));
#pragma unroll
    for (int g = 0; g < CHUNK_Z / PB; ++g)
    {
	const int zb = cl.z0 + g * PB;
	bool a0[PB], a1[PB];
	unsigned m0[PB], m1[PB];
#pragma unroll
	for (int q = 0; q < PB; ++q)
	{
	    const uchar2 l = cl.lab[g * PB + q];
	    a0[q] = (l.x == L_INTERIOR);
	    a1[q] = (l.y == L_INTERIOR);
	    const bool act = a0[q] | a1[q];
	    m0[q] = !act ? 0u : (MODE == SM_JACOBI_ZERO ? unsigned(cl.flg[g * PB + q].x) : 0x7fu);
	    m1[q] = !act ? 0u : (MODE == SM_JACOBI_ZERO ? unsigned(cl.flg[g * PB + q].y) : 0x7fu);
	}
	T2 c[PB + 2];
#pragma unroll
	for (int k = 0; k < PB + 2; ++k)
	{
	    unsigned need = 0u;
	    if (k >= 1 && k <= PB) need |= (m0[k - 1] | m1[k - 1]) & 1u;   // centre of plane k - 1
	    if (k <= PB - 1) need |= (m0[k] | m1[k]) & 32u;                // -z neighbour of plane k
	    if (k >= 2) need |= (m0[k - 2] | m1[k - 2]) & 64u;            // +z neighbour of plane k - 2
	    c[k] = need ? ld2(a.in + int64_t(zb - 1 + k) * a.plane + inPlane) : zero2;
	}
	T xm[PB], xp[PB];
	T2 ym[PB], yp[PB], rhs[PB];
#pragma unroll
	for (int q = 0; q < PB; ++q)
	{
	    const int64_t i = int64_t(zb + q) * a.plane + inPlane;
	    const unsigned m = m0[q] | m1[q];
	    xm[q] = (m0[q] & 2u) ? a.in[i - 1] : T(0.0);
	    xp[q] = (m1[q] & 4u) ? a.in[i + 2] : T(0.0);
	    ym[q] = (m & 8u) ? ld2(a.in + i - a.pitch) : zero2;
	    yp[q] = (m & 16u) ? ld2(a.in + i + a.pitch) : zero2;
	    rhs[q] = (MODE != SM_APPLY && (a0[q] | a1[q])) ? ld2(a.b + i) : zero2;
	}
#pragma unroll
	for (int q = 0; q < PB; ++q)
	{
	    if (!(a0[q] | a1[q])) continue;
	    const int z = zb + q;
	    const int64_t i = int64_t(z) * a.plane + inPlane;
	    T2 c2 = c[q + 1], zm = c[q], zp = c[q + 2], ymq = ym[q], ypq = yp[q];
	    if (MODE == SM_JACOBI_ZERO)
	    {
		if (!(m0[q] & 1u)) c2.x = T(0.0);
		if (!(m1[q] & 1u)) c2.y = T(0.0);
		if (!(m0[q] & 8u)) ymq.x = T(0.0);
		if (!(m1[q] & 8u)) ymq.y = T(0.0);
		if (!(m0[q] & 16u)) ypq.x = T(0.0);
		if (!(m1[q] & 16u)) ypq.y = T(0.0);
		if (!(m0[q] & 32u)) zm.x = T(0.0);
		if (!(m1[q] & 32u)) zm.y = T(0.0);
		if (!(m0[q] & 64u)) zp.x = T(0.0);
		if (!(m1[q] & 64u)) zp.y = T(0.0);
	    }
	    T lap0 = -xm[q];
	    lap0 -= c2.y; lap0 -= ymq.x; lap0 -= ypq.x; lap0 -= zm.x; lap0 -= zp.x;
	    lap0 += T(6.0) * c2.x;
	    T lap1 = -c2.x;
	    lap1 -= xp[q]; lap1 -= ymq.y; lap1 -= ypq.y; lap1 -= zm.y; lap1 -= zp.y;
	    lap1 += T(6.0) * c2.y;
	    const T o0 = stencilFinish<MODE, T>(lap0, c2.x, rhs[q].x, T(6.0));
	    const T o1 = stencilFinish<MODE, T>(lap1, c2.y, rhs[q].y, T(6.0));
	    if (a0[q] & a1[q]) st2(a.out + i, make2<T>(o0, o1));
	    else if (a0[q]) a.out[i] = o0;
	    else a.out[i + 1] = o1;
	    if (DOT && z >= a.dotLo && z < a.dotHi) acc += double((a0[q] ? c2.x * lap0 : T(0.0)) + (a1[q] ? c2.y * lap1 : T(0.0)));
	}
    }
    return acc;
}

template <typename T, int MODE, bool DOT>
__device__ __forceinline__ double stencilBody(const StencilArgsT<T> &a, int vb, int tid)
{
    double acc = 0.0;
    if (vb < a.nChunks)
    {
	ChunkLabels cl;
	stencilLabels<T, MODE>(a, vb, tid, cl);
	pdlWait();
	acc = stencilChunk<T, MODE, DOT>(a, cl);
    }
    else acc = stencilBoundary<T, MODE, DOT>(a, (vb - a.nChunks) * BLOCK + tid);
    return acc;
}

// MINB = 6 trades loads in flight per thread for resident CTAs (56 -> 40 registers): see launchStencil
template <int MODE, bool DOT, typename T = double, int MINB = 1>
__global__ void __launch_bounds__(BLOCK, MINB) k_stencil(const StencilArgsT<T> a)
{
    pdlLaunch();
    const double acc = stencilBody<T, MODE, DOT>(a, blockIdx.x, threadIdx.x);
    if (DOT) gridReduce(acc, a.partials, a.ticket, a.result);
}

// k_stencil with stencilChunkBatched: PB = 4 keeps a whole chunk's values in registers (2 CTAs per SM), PB = 2 half of it (3 CTAs per SM)
template <int MODE, bool DOT, int PB>
__global__ void __launch_bounds__(BLOCK, PB >= 4 ? 2 : 3) k_stencil_b(const StencilArgsT<double> a)
{
    pdlLaunch();
    double acc = 0.0;
    const int vb = blockIdx.x, tid = threadIdx.x;
    if (vb < a.nChunks)
    {
	ChunkLabels cl;
	stencilLabels<double, MODE>(a, vb, tid, cl);
	pdlWait();
	acc = stencilChunkBatched<double, MODE, DOT, PB>(a, cl);
    }
    else acc = stencilBoundary<double, MODE, DOT>(a, (vb - a.nChunks) * BLOCK + tid);
    if (DOT) gridReduce(acc, a.partials, a.ticket, a.result);
}

// The same operator as a PERSISTENT kernel: the grid is what fits the device at once, a CTA walks the virtual CTAs vb, vb + grid, ...
// and loads the chunk id and labels of its NEXT chunk before it computes the current one.  A virtual CTA of k_stencil is a chain of
// three dependent round trips (chunk id -> labels -> values); only the first wave overlaps the first two with its predecessor (PDL),
// and a 256^3 level is 4-6 waves.  Here every chunk after a CTA's first costs one round trip.  Same per-cell arithmetic; the fused dot
// product sums over another decomposition (different rounding of p.Ap, like the TMA kernel).
template <int MODE, bool DOT, typename T = double>
__global__ void __launch_bounds__(BLOCK, 6) k_stencil_loop(const StencilArgsT<T> a, int nVirtual)
{
    pdlLaunch();
    const int tid = threadIdx.x;
    double acc = 0.0;
    int vb = blockIdx.x;
    ChunkLabels cur = {};
    if (vb < a.nChunks) stencilLabels<T, MODE>(a, vb, tid, cur);
    pdlWait();
    while (vb < a.nChunks)
    {
	const int nx = vb + int(gridDim.x);
	ChunkLabels nxt = {};
	if (nx < a.nChunks) stencilLabels<T, MODE>(a, nx, tid, nxt);
	acc += stencilChunk<T, MODE, DOT>(a, cur);
	cur = nxt;
	vb = nx;
    }
    for (; vb < nVirtual; vb += int(gridDim.x)) acc += stencilBoundary<T, MODE, DOT>(a, (vb - a.nChunks) * BLOCK + tid);
    if (DOT) gridReduce(acc, a.partials, a.ticket, a.result);
}


// ------------------------------------------------------------------------------------------------
// The same operator with the input staged through shared memory by the Tensor Memory Accelerator (the A/B SURVEY.md
// section 7 step 7 asks for).  One CTA owns a 64 x 8 x 4 brick of storage cells; ONE elected thread issues a 3D
// cp.async.bulk.tensor load of the brick plus its halo -- a 68 x 10 x 6 box of `in`, starting two cells left of the brick so
// that the aligned x-pairs stay 16-byte aligned in shared memory -- and the CTA waits on an mbarrier for the bytes to land.
// Out-of-range coordinates (the box reaches past the storage box at every edge brick) are zero-filled by the TMA unit, which
// is exactly the value a vector grid has there.  Every stencil read then comes from shared memory: the seven global loads
// per thread and plane of k_stencil (five of them re-reading lines a neighbouring thread or plane also reads, through
// L1/L2) become one bulk copy of each line per brick.  Right-hand side and labels are read once per cell anyway: plain
// coalesced loads, issued before the wait.  Arithmetic and its order are k_stencil's: bitwise the same result.
// CTAs [0, nBricks): bricks holding an INTERIOR cell; the rest: the boundary records, exactly as in k_stencil.
// ------------------------------------------------------------------------------------------------
constexpr int TB_X = 64, TB_Y = 8, TB_Z = 4;
constexpr int TB_BOX_X = TB_X + 4, TB_BOX_Y = TB_Y + 2, TB_BOX_Z = TB_Z + 2;
constexpr unsigned TB_BOX_BYTES = TB_BOX_X * TB_BOX_Y * TB_BOX_Z * sizeof(double);

__device__ __forceinline__ unsigned smemAddr(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long *bar, unsigned phase)
{
    asm volatile(
	"{\n"
	".reg .pred p;\n"
	"WAIT_LOOP:\n"
	"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	"@p bra WAIT_DONE;\n"
	"bra WAIT_LOOP;\n"
	"WAIT_DONE:\n"
	"}\n" ::"r"(smemAddr(bar)),
	"r"(phase)
	: "memory");
}
__device__ __forceinline__ void tmaLoad3d(void *dst, const void *tensorMap, int c0, int c1, int c2, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smemAddr(dst)),
		 "l"(tensorMap), "r"(c0), "r"(c1), "r"(c2), "r"(smemAddr(bar))
		 : "memory");
}

template <int MODE, bool DOT>
__global__ void __launch_bounds__(BLOCK) k_stencil_tma(const StencilArgs a, const __grid_constant__ TmaMap tmIn, const int32_t *bricks, int nBricks,
						      int bricksX, int bricksY, int ny)
{
    pdlLaunch();
    __shared__ alignas(128) double tile[TB_BOX_Z][TB_BOX_Y][TB_BOX_X];
    __shared__ alignas(8) unsigned long long bar;
    double acc = 0.0;
    if (int(blockIdx.x) < nBricks)
    {
	const int br = bricks[blockIdx.x];
	const int bz = br / (bricksX * bricksY), by = (br - bz * bricksX * bricksY) / bricksX, bx = br - (bz * bricksY + by) * bricksX;
	const int x0 = bx * TB_X, y0 = by * TB_Y, z0 = bz * TB_Z;
	const int px = threadIdx.x & 31, ty = threadIdx.x >> 5;
	const int x = x0 + 2 * px, y = y0 + ty;
	const bool inRow = x < a.pitch && y < ny;
	const int64_t inPlane = int64_t(y) * a.pitch + x;
	// prologue: labels of the thread's cells (static)
	uchar2 lab[TB_Z];
#pragma unroll
	for (int dz = 0; dz < TB_Z; ++dz)
	{
	    const int z = z0 + dz;
	    lab[dz] = make_uchar2(L_EXTERIOR, L_EXTERIOR);
	    if (inRow && z >= a.zlo && z < a.zhi) lab[dz] = *reinterpret_cast<const uchar2 *>(a.labels + int64_t(z) * a.plane + inPlane);
	}
	if (threadIdx.x == 0) mbarInit(&bar, 1);
	__syncthreads();
	pdlWait();
	if (threadIdx.x == 0)
	{
	    mbarExpectTx(&bar, TB_BOX_BYTES);
	    // tensor coordinates: x, y, and z + 1 (the tensor starts at the grid's lower guard plane)
	    tmaLoad3d(&tile[0][0][0], &tmIn, x0 - 2, y0 - 1, z0 - 1 + 1, &bar);
	}
	double2 rhs[TB_Z];
#pragma unroll
	for (int dz = 0; dz < TB_Z; ++dz)
	{
	    rhs[dz] = make_double2(0.0, 0.0);
	    const bool any = (lab[dz].x == L_INTERIOR) | (lab[dz].y == L_INTERIOR);
	    if (MODE != SM_APPLY && any) rhs[dz] = ld2(a.b + int64_t(z0 + dz) * a.plane + inPlane);
	}
	mbarWait(&bar, 0);
	const int cx = 2 * px + 2, cy = ty + 1;
	double2 zm = *reinterpret_cast<const double2 *>(&tile[0][cy][cx]);
	double2 c2 = *reinterpret_cast<const double2 *>(&tile[1][cy][cx]);
#pragma unroll
	for (int dz = 0; dz < TB_Z; ++dz)
	{
	    const int z = z0 + dz;
	    const double2 zp = *reinterpret_cast<const double2 *>(&tile[dz + 2][cy][cx]);
	    const uchar2 l = lab[dz];
	    const bool a0 = (l.x == L_INTERIOR), a1 = (l.y == L_INTERIOR);
	    if (a0 | a1)
	    {
		const int64_t i = int64_t(z) * a.plane + inPlane;
		const double xm = tile[dz + 1][cy][cx - 1], xp = tile[dz + 1][cy][cx + 2];
		const double2 ym = *reinterpret_cast<const double2 *>(&tile[dz + 1][cy - 1][cx]);
		const double2 yp = *reinterpret_cast<const double2 *>(&tile[dz + 1][cy + 1][cx]);
		double lap0 = -xm;
		lap0 -= c2.y; lap0 -= ym.x; lap0 -= yp.x; lap0 -= zm.x; lap0 -= zp.x;
		lap0 += 6.0 * c2.x;
		double lap1 = -c2.x;
		lap1 -= xp; lap1 -= ym.y; lap1 -= yp.y; lap1 -= zm.y; lap1 -= zp.y;
		lap1 += 6.0 * c2.y;
		const double o0 = stencilFinish<MODE, double>(lap0, c2.x, rhs[dz].x, 6.0);
		const double o1 = stencilFinish<MODE, double>(lap1, c2.y, rhs[dz].y, 6.0);
		if (a0 & a1) st2(a.out + i, make_double2(o0, o1));
		else if (a0) a.out[i] = o0;
		else a.out[i + 1] = o1;
		if (DOT && z >= a.dotLo && z < a.dotHi) acc += (a0 ? c2.x * lap0 : 0.0) + (a1 ? c2.y * lap1 : 0.0);
	    }
	    zm = c2;
	    c2 = zp;
	}
    }
    else acc = stencilBoundary<double, MODE, DOT>(a, (int(blockIdx.x) - nBricks) * BLOCK + int(threadIdx.x));
    if (DOT) gridReduce(acc, a.partials, a.ticket, a.result);
}

// brick flags: 1 when the 64 x 8 x 4 brick holds an INTERIOR cell
__global__ void __launch_bounds__(BLOCK) k_brick_flags(uint8_t *flags, const uint8_t *labels, int bricksX, int bricksY, int pitch, int64_t plane, int ny, int nz)
{
    const int br = blockIdx.x;
    const int bz = br / (bricksX * bricksY), by = (br - bz * bricksX * bricksY) / bricksX, bx = br - (bz * bricksY + by) * bricksX;
    const int x = bx * TB_X + 2 * (threadIdx.x & 31), y = by * TB_Y + (threadIdx.x >> 5);
    int any = 0;
    if (x < pitch && y < ny)
	for (int dz = 0; dz < TB_Z; ++dz)
	{
	    const int z = bz * TB_Z + dz;
	    if (z >= nz) break;
	    const uchar2 l = *reinterpret_cast<const uchar2 *>(labels + int64_t(z) * plane + int64_t(y) * pitch + x);
	    any |= (l.x == L_INTERIOR) | (l.y == L_INTERIOR);
	}
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) flags[br] = uint8_t(any);
}

// ------------------------------------------------------------------------------------------------
// Boundary-band damped Jacobi (Ops.h:524-619).  The reference computes every band cell from the
// current grid into a temporary list, then writes the list back.  Here sweep s reads band
// neighbours from compact array vin (the state after sweep s-1) and frozen non-band neighbours
// from the grid, writing compact array vout -- or the grid itself on the last sweep, which is
// race-free because nobody reads band cells from the grid in that sweep.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct BandArgsT
{
    T *x;                // grid
    const T *b;          // grid rhs
    const int32_t *bandIdx;
    const int32_t *bandRef;  // [6][nBand] neighbour reference: >= 0 position in the band; BAND_SKIP: coefficient 0 (not active);
			     // <= -2: an active cell outside the band (frozen during the band sweeps) at grid index -2 - ref
    const double *bcoef;
    const unsigned short *wcode;  // coefficient codes of the BOUNDARY cells (k_band_coef)
    const T *vin;
    T *vout;
    T *bandB;
    int nBoundary, nBand;
    int pitch;
    int64_t plane;
};
typedef BandArgsT<double> BandArgs;
constexpr int BAND_SKIP = -1;
constexpr int BAND_PER_THREAD = 2;  // cells per thread: two independent gather chains in flight

// FROM_COMPACT: centre/band-neighbour values come from vin; TO_GRID: result goes to x[idx];
// FIRST: rhs is gathered from the grid and cached in bandB; ZERO: the grid is known to be all zero;
// HAS_W: level 0 with face weights -- BOUNDARY cells multiply by their coefficient records (elsewhere every coefficient is 1)
// FZ: the grid was all zero when this sweep group started, so the frozen (non-band) neighbours are zero and are not read at
// all -- the grid may hold anything off the band (no zero fill before the group, see SM_JACOBI_ZERO)
// Prologue (static metadata, before pdlWait): grid index, neighbour references, diagonal, coefficient record.
// CGL: values of the compact arrays are read past L1 (ld.cg) -- inside the persistent sweep-group kernel another SM wrote them
// during the same launch
template <bool FROM_COMPACT, bool TO_GRID, bool FIRST, bool ZERO, bool HAS_W, bool FZ = false, bool CGL = false, typename T = double, int BAND_PER_THREAD = gmg::BAND_PER_THREAD>
__device__ __forceinline__ void bandBody(const BandArgsT<T> &a, int vb, int tid)
{
    int64_t gi[BAND_PER_THREAD];
    int ref[BAND_PER_THREAD][6];
    T diag[BAND_PER_THREAD];
    unsigned code[BAND_PER_THREAD];
#pragma unroll
    for (int c = 0; c < BAND_PER_THREAD; ++c)
    {
	const int k = (vb * BAND_PER_THREAD + c) * BLOCK + tid;
	gi[c] = 0;
	diag[c] = T(6.0);
	code[c] = 0x555u;  // every coefficient 1
#pragma unroll
	for (int n = 0; n < 6; ++n) ref[c][n] = BAND_SKIP;
	if (k >= a.nBand) continue;
	if (!FROM_COMPACT || TO_GRID || FIRST) gi[c] = a.bandIdx[k];
	if (k < a.nBoundary) diag[c] = T(a.bcoef[int64_t(6) * a.nBoundary + k]);
	if (HAS_W && !ZERO && k < a.nBoundary) code[c] = a.wcode[k];
	if (!ZERO)
	{
#pragma unroll
	    for (int n = 0; n < 6; ++n) ref[c][n] = a.bandRef[int64_t(n) * a.nBand + k];
	}
    }
    pdlWait();
    T v[BAND_PER_THREAD];
#pragma unroll
    for (int c = 0; c < BAND_PER_THREAD; ++c)
    {
	const int k = (vb * BAND_PER_THREAD + c) * BLOCK + tid;
	v[c] = T(0.0);
	if (k >= a.nBand) continue;
	const int64_t i = gi[c];
	T rhs;
	if (FIRST) { rhs = a.b[i]; a.bandB[k] = rhs; }
	else rhs = a.bandB[k];
	T centre = T(0.0), lap = T(0.0);
	if (!ZERO)
	{
	    const bool weighted = HAS_W && k < a.nBoundary;
	    centre = FROM_COMPACT ? (CGL ? __ldcg(a.vin + k) : a.vin[k]) : a.x[i];
	    const int64_t stride[6] = {-1, 1, -int64_t(a.pitch), int64_t(a.pitch), -a.plane, a.plane};
#pragma unroll
	    for (int n = 0; n < 6; ++n)
	    {
		const int r = ref[c][n];
		if (r == BAND_SKIP) continue;
		T u;
		if (FROM_COMPACT && FZ)
		{
		    if (r < 0) continue;  // a frozen neighbour holds 0: lap -= c * 0 changes nothing
		    u = CGL ? __ldcg(a.vin + r) : a.vin[r];
		}
		else if (FROM_COMPACT) u = r >= 0 ? (CGL ? __ldcg(a.vin + r) : a.vin[r]) : a.x[-2 - r];
		else u = a.x[i + stride[n]];
		if (weighted)
		{
		    const unsigned cc = (code[c] >> (2 * n)) & 3u;
		    if (cc == 1u) lap -= u;  // 1 * u, bit for bit
		    else if (cc == 2u) lap -= T(a.bcoef[int64_t(n) * a.nBoundary + k]) * u;  // fractional: rare
		}
		else lap -= u;
	    }
	    lap += diag[c] * centre;
	}
	T r = rhs - lap;
	r /= diag[c];
	v[c] = centre + T(2.0 / 3.0) * r;
    }
#pragma unroll
    for (int c = 0; c < BAND_PER_THREAD; ++c)
    {
	const int k = (vb * BAND_PER_THREAD + c) * BLOCK + tid;
	if (k >= a.nBand) continue;
	if (TO_GRID) a.x[gi[c]] = v[c];
	else a.vout[k] = v[c];
    }
}
// PT: cells per thread.  A sweep is one dependent round of gathers per CTA, so a grid that does not fit the device at once pays the
// round twice (256^3, level 0: 1211 CTAs of 2 cells per thread over 888 resident slots); three cells per thread bring it back to
// one wave (launchBand picks)
template <bool FROM_COMPACT, bool TO_GRID, bool FIRST, bool ZERO, bool HAS_W, bool FZ = false, typename T = double, int PT = BAND_PER_THREAD>
__global__ void __launch_bounds__(BLOCK, PT > 2 ? 6 : 1) k_band(const BandArgsT<T> a)
{
    pdlLaunch();
    bandBody<FROM_COMPACT, TO_GRID, FIRST, ZERO, HAS_W, FZ, false, T, PT>(a, blockIdx.x, threadIdx.x);
}


// ------------------------------------------------------------------------------------------------
// A whole group of band sweeps (MG.cpp:445-513: myBoundarySmootherIterations = 3 before and after every interior sweep) in ONE
// launch: the sweeps of a group are dependent steps of a few microseconds each, and a grid-wide barrier between them (one
// atomic per CTA on a counter, everybody polls a generation word) is cheaper than a kernel boundary with its drain, launch and
// ramp.  The grid is sized to be co-resident (the host asks the occupancy calculator); CTAs walk the virtual CTAs of the
// sweep-per-launch kernels with the same bodies, so every cell sees the same arithmetic.  A wait that exceeds two seconds of
// wall clock sets the error word and gives up instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------
struct GroupBarrier
{
    unsigned count;
    unsigned generation;
    int error;
};
__device__ __forceinline__ void gridBarrier(GroupBarrier *bar)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
	__threadfence();
	const unsigned gen = *reinterpret_cast<volatile unsigned *>(&bar->generation);
	if (atomicAdd(&bar->count, 1u) == gridDim.x - 1)
	{
	    bar->count = 0u;
	    __threadfence();
	    atomicAdd(&bar->generation, 1u);
	}
	else
	{
	    unsigned long long t0 = 0;
	    unsigned spins = 0;
	    while (*reinterpret_cast<volatile unsigned *>(&bar->generation) == gen)
	    {
		if ((++spins & 4095u) == 0)
		{
		    unsigned long long t;
		    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		    if (t0 == 0) t0 = t;
		    else if (t - t0 > 2000000000ull) { atomicExch(&bar->error, 1); break; }
		}
	    }
	}
	__threadfence();
    }
    __syncthreads();
}

template <bool HAS_W, bool ZEROGRID>
__global__ void __launch_bounds__(BLOCK) k_band_group(BandArgs a, int sweeps, int nVirtual, GroupBarrier *bar)
{
    pdlLaunch();
    double *cur = a.vout, *nxt = const_cast<double *>(a.vin);  // the two compact arrays: sweep 1 writes cur
    // sweep 1: grid -> compact
    a.vin = nullptr;
    a.vout = cur;
    for (int vb = blockIdx.x; vb < nVirtual; vb += gridDim.x)
    {
	if (ZEROGRID) bandBody<false, false, true, true, false>(a, vb, threadIdx.x);
	else bandBody<false, false, true, false, HAS_W>(a, vb, threadIdx.x);
    }
    for (int sw = 2; sw <= sweeps; ++sw)
    {
	gridBarrier(bar);
	a.vin = cur;
	a.vout = nxt;
	for (int vb = blockIdx.x; vb < nVirtual; vb += gridDim.x)
	{
	    if (sw == sweeps) bandBody<true, true, false, false, HAS_W, ZEROGRID, true>(a, vb, threadIdx.x);
	    else bandBody<true, false, false, false, HAS_W, ZEROGRID, true>(a, vb, threadIdx.x);
	}
	double *t = cur; cur = nxt; nxt = t;
    }
}


// ------------------------------------------------------------------------------------------------
// The sweep group with its per-cell metadata RESIDENT for the whole group.  k_band_group above only replaces
// kernel boundaries by barriers -- over ~900 CTAs, measured slower.  Here the grid is two 512-thread CTAs per SM (296 CTAs: a
// barrier costs ~1.7 us, scripts/barrier_probe.cu), every thread OWNS up to CPT band cells for all sweeps of the group, and
// what a sweep needs besides the neighbours' values stays on chip: right-hand side, diagonal, current value and coefficient
// code in registers, the six neighbour references in shared memory (24 B per cell: 148 SMs x 2 x 512 x CPT cells of capacity).
// Sweep 1 reads everything once; sweeps 2.. only gather the neighbours' values (compact array, ld.cg: written by other SMs
// during this launch) and the few frozen / fractional ones.  Per cell the arithmetic and its order are bandBody's.
// ------------------------------------------------------------------------------------------------
constexpr int BG_THREADS = 512;

template <bool HAS_W, bool ZEROGRID, int CPT>
__global__ void __launch_bounds__(BG_THREADS, 2) k_band_resident(BandArgs a, int sweeps, GroupBarrier *bar)
{
    pdlLaunch();
    extern __shared__ int refs[];  // [CPT][6][BG_THREADS]
    const int tid = threadIdx.x;
    const int gthreads = gridDim.x * BG_THREADS, gtid = blockIdx.x * BG_THREADS + tid;
    double rhs[CPT], diag[CPT], val[CPT];
    unsigned code[CPT];
    // prologue (static): references into shared memory, diagonal, coefficient code (the grid index is re-read where it is
    // needed, first and last sweep: registers are what limits the cells a thread can own)
#pragma unroll
    for (int c = 0; c < CPT; ++c)
    {
	const int k = c * gthreads + gtid;
	diag[c] = 6.0;
	code[c] = 0x555u;
	if (k >= a.nBand) continue;
#pragma unroll
	for (int n = 0; n < 6; ++n) refs[(c * 6 + n) * BG_THREADS + tid] = a.bandRef[int64_t(n) * a.nBand + k];
	if (k < a.nBoundary)
	{
	    diag[c] = a.bcoef[int64_t(6) * a.nBoundary + k];
	    if (HAS_W) code[c] = a.wcode[k];
	}
    }
    pdlWait();
    double *cur = a.vout, *nxt = const_cast<double *>(a.vin);
    // sweep 1: grid -> compact (bandBody<false, false, true, ZEROGRID, HAS_W>)
#pragma unroll
    for (int c = 0; c < CPT; ++c)
    {
	const int k = c * gthreads + gtid;
	if (k >= a.nBand) continue;
	const int64_t i = a.bandIdx[k];
	rhs[c] = a.b[i];
	double centre = 0.0, lap = 0.0;
	if (!ZEROGRID)
	{
	    const bool weighted = HAS_W && k < a.nBoundary;
	    centre = a.x[i];
	    const int64_t stride[6] = {-1, 1, -int64_t(a.pitch), int64_t(a.pitch), -a.plane, a.plane};
#pragma unroll
	    for (int n = 0; n < 6; ++n)
	    {
		if (refs[(c * 6 + n) * BG_THREADS + tid] == BAND_SKIP) continue;
		const double u = a.x[i + stride[n]];
		if (weighted)
		{
		    const unsigned cc = (code[c] >> (2 * n)) & 3u;
		    if (cc == 1u) lap -= u;
		    else if (cc == 2u) lap -= a.bcoef[int64_t(n) * a.nBoundary + k] * u;
		}
		else lap -= u;
	    }
	    lap += diag[c] * centre;
	}
	double r = rhs[c] - lap;
	r /= diag[c];
	val[c] = centre + (2.0 / 3.0) * r;
	if (sweeps == 1) a.x[i] = val[c];
	else cur[k] = val[c];
    }
    for (int sw = 2; sw <= sweeps; ++sw)
    {
	gridBarrier(bar);
	// compact -> compact, the last one compact -> grid (bandBody<true, last, false, false, HAS_W, ZEROGRID, true>)
#pragma unroll
	for (int c = 0; c < CPT; ++c)
	{
	    const int k = c * gthreads + gtid;
	    if (k >= a.nBand) continue;
	    const bool weighted = HAS_W && k < a.nBoundary;
	    const double centre = val[c];
	    double lap = 0.0;
#pragma unroll
	    for (int n = 0; n < 6; ++n)
	    {
		const int r = refs[(c * 6 + n) * BG_THREADS + tid];
		if (r == BAND_SKIP) continue;
		double u;
		if (r >= 0) u = __ldcg(cur + r);
		else if (ZEROGRID) continue;  // a frozen neighbour holds 0
		else u = a.x[-2 - r];
		if (weighted)
		{
		    const unsigned cc = (code[c] >> (2 * n)) & 3u;
		    if (cc == 1u) lap -= u;
		    else if (cc == 2u) lap -= a.bcoef[int64_t(n) * a.nBoundary + k] * u;
		}
		else lap -= u;
	    }
	    lap += diag[c] * centre;
	    double rr = rhs[c] - lap;
	    rr /= diag[c];
	    val[c] = centre + (2.0 / 3.0) * rr;
	    if (sw == sweeps) a.x[a.bandIdx[k]] = val[c];
	    else nxt[k] = val[c];
	}
	double *t = cur; cur = nxt; nxt = t;
    }
}

template <typename T = double>
__global__ void __launch_bounds__(BLOCK) k_band_scatter(T *x, const int32_t *bandIdx, const T *v, int nBand)
{
    pdlLaunch();
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    const int i = k < nBand ? bandIdx[k] : 0;
    pdlWait();
    if (k < nBand) x[i] = v[k];
}

// ------------------------------------------------------------------------------------------------
// Restriction (Ops.h:734-835): coarse (active) = sum_{z,y,x} w[x]w[y]w[z] fine(2c-1+(x,y,z)), weights (1,3,3,1)/8.
// One thread per coarse cell over the chunks holding active coarse cells.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct TransferArgsT
{
    const uint8_t *fineLabels, *coarseLabels;
    const T *fine;
    const T *coarse;
    T *out;
    const int32_t *chunks;
    int chunksPerPlane;
    int finePitch, coarsePitch;
    int64_t finePlane, coarsePlane;
    int fineNz, coarseNz, coarseNy;
    int shift[3];
    int zlo, zhi;  // clip on the destination's z-planes (coarse planes for restriction, fine planes for prolongation)
};
typedef TransferArgsT<double> TransferArgs;

// One virtual CTA covers one z-plane and one 256-cell half of a chunk (8 virtual CTAs per chunk): a coarse cell is 64 fine
// reads, so the work is spread over as many threads as there are coarse cells.
constexpr int RESTRICT_SPLIT = CHUNK_Z * 2;
template <typename T>
__device__ __forceinline__ void restrictBody(const TransferArgsT<T> &a, int vb, int tid)
{
    const T rw[4] = {T(1. / 8.), T(3. / 8.), T(3. / 8.), T(1. / 8.)};
    const int c = a.chunks[vb / RESTRICT_SPLIT];
    const int sub = vb % RESTRICT_SPLIT;
    const int zb = c / a.chunksPerPlane;
    const int64_t base = int64_t(c - zb * a.chunksPerPlane) * CHUNK_CELLS;
    const int cz = zb * CHUNK_Z + (sub >> 1);
    if (cz >= a.zhi || cz < a.zlo) return;
    const int64_t inPlane = base + (sub & 1) * BLOCK + tid;
    if (inPlane >= a.coarsePlane) return;
    const int64_t ci = int64_t(cz) * a.coarsePlane + inPlane;
    const int l = a.coarseLabels[ci];
    pdlWait();
    if (!(l == L_INTERIOR || l == L_BOUNDARY)) return;
    const int cy = int(unsigned(inPlane) / unsigned(a.coarsePitch)), cx = int(inPlane - int64_t(cy) * a.coarsePitch);
    const int fx = 2 * (cx - a.shift[0]) - 1, fy = 2 * (cy - a.shift[1]) - 1, fz = 2 * (cz - a.shift[2]) - 1;
    const T *f = a.fine + (int64_t(fz) * a.finePlane + int64_t(fy) * a.finePitch + fx);
    T v = T(0.0);
#pragma unroll
    for (int z = 0; z < 4; ++z)
#pragma unroll
	for (int y = 0; y < 4; ++y)
	{
	    const T *row = f + int64_t(z) * a.finePlane + int64_t(y) * a.finePitch;
	    // fx is odd: row[1..2] is an aligned pair
	    const T s0 = row[0];
	    const typename Vec2<T>::type s12 = ld2(row + 1);
	    const T s3 = row[3];
	    v += rw[0] * rw[y] * rw[z] * s0;
	    v += rw[1] * rw[y] * rw[z] * s12.x;
	    v += rw[2] * rw[y] * rw[z] * s12.y;
	    v += rw[3] * rw[y] * rw[z] * s3;
	}
    a.out[ci] = v;
}
template <typename T = double>
__global__ void __launch_bounds__(BLOCK) k_restrict(const TransferArgsT<T> a) { pdlLaunch(); restrictBody<T>(a, blockIdx.x, threadIdx.x); }


// ------------------------------------------------------------------------------------------------
// Restriction with the fine residual staged through shared memory by TMA.  k_restrict issues 48 global loads per coarse
// cell (16 rows x {scalar, aligned pair, scalar}); every fine value is wanted by 8 coarse cells, so the load/store unit
// sees 8x the data and the kernel sits at 0.40 of the HBM roofline with its warps waiting on L1/L2 (ncu: 412 M sectors
// at 512^3 against 45-60 M for the stencil kernels).  Here a CTA owns a 32 x 4 x 2 brick of coarse cells; its 4x4x4 taps
// cover a 66 x 10 x 6 box of fine cells, fetched by ONE bulk tensor copy (68 wide, one cell further left, so the aligned
// pair of every row is 16-byte aligned in shared memory -- the same box shape as k_stencil_tma, hence the same tensor
// map), and the 48 loads per coarse cell become shared-memory loads.  Same arithmetic, same order: bitwise k_restrict.
// ------------------------------------------------------------------------------------------------
constexpr int RB_X = 32, RB_Y = 4, RB_Z = 2;

__global__ void __launch_bounds__(BLOCK) k_restrict_tma(const TransferArgs a, const __grid_constant__ TmaMap tmFine, const int32_t *bricks, int bricksX, int bricksY)
{
    pdlLaunch();
    __shared__ alignas(128) double tile[TB_BOX_Z][TB_BOX_Y][TB_BOX_X];
    __shared__ alignas(8) unsigned long long bar;
    const int br = bricks[blockIdx.x];
    const int bz = br / (bricksX * bricksY), by = (br - bz * bricksX * bricksY) / bricksX, bx = br - (bz * bricksY + by) * bricksX;
    const int c0x = bx * RB_X, c0y = by * RB_Y, c0z = bz * RB_Z;
    const int tx = threadIdx.x & 31, ty = (threadIdx.x >> 5) & 3, tz = threadIdx.x >> 7;
    const int cx = c0x + tx, cy = c0y + ty, cz = c0z + tz;
    const bool inBox = cx < a.coarsePitch && cy < a.coarseNy && cz >= a.zlo && cz < a.zhi;
    const int64_t ci = int64_t(cz) * a.coarsePlane + int64_t(cy) * a.coarsePitch + cx;
    const int l = inBox ? int(a.coarseLabels[ci]) : int(L_EXTERIOR);  // prologue: static
    if (threadIdx.x == 0) mbarInit(&bar, 1);
    __syncthreads();
    pdlWait();
    if (threadIdx.x == 0)
    {
	mbarExpectTx(&bar, TB_BOX_BYTES);
	// first tap of the brick's first coarse cell: fine 2 (c - shift) - 1; the box starts one cell further left in x; z + 1 = guard plane
	tmaLoad3d(&tile[0][0][0], &tmFine, 2 * (c0x - a.shift[0]) - 2, 2 * (c0y - a.shift[1]) - 1, 2 * (c0z - a.shift[2]) - 1 + 1, &bar);
    }
    mbarWait(&bar, 0);
    if (!(l == L_INTERIOR || l == L_BOUNDARY)) return;
    const double rw[4] = {1. / 8., 3. / 8., 3. / 8., 1. / 8.};
    double v = 0.0;
#pragma unroll
    for (int z = 0; z < 4; ++z)
#pragma unroll
	for (int y = 0; y < 4; ++y)
	{
	    const double *row = &tile[2 * tz + z][2 * ty + y][2 * tx + 1];
	    const double s0 = row[0];
	    const double2 s12 = *reinterpret_cast<const double2 *>(row + 1);
	    const double s3 = row[3];
	    v += rw[0] * rw[y] * rw[z] * s0;
	    v += rw[1] * rw[y] * rw[z] * s12.x;
	    v += rw[2] * rw[y] * rw[z] * s12.y;
	    v += rw[3] * rw[y] * rw[z] * s3;
	}
    a.out[ci] = v;
}

// coarse-brick flags: 1 when the 32 x 4 x 2 brick of coarse cells holds an active cell
__global__ void __launch_bounds__(BLOCK) k_cbrick_flags(uint8_t *flags, const uint8_t *labels, int bricksX, int bricksY, int pitch, int64_t plane, int ny, int nz)
{
    const int br = blockIdx.x;
    const int bz = br / (bricksX * bricksY), by = (br - bz * bricksX * bricksY) / bricksX, bx = br - (bz * bricksY + by) * bricksX;
    const int x = bx * RB_X + (threadIdx.x & 31), y = by * RB_Y + ((threadIdx.x >> 5) & 3), z = bz * RB_Z + (threadIdx.x >> 7);
    int any = 0;
    if (x < pitch && y < ny && z < nz)
    {
	const int l = labels[int64_t(z) * plane + int64_t(y) * pitch + x];
	any = (l == L_INTERIOR || l == L_BOUNDARY);
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) flags[br] = uint8_t(any);
}

// ------------------------------------------------------------------------------------------------
// Prolongation (Ops.h:873-972): fine (active) += 4 * trilerp(8 coarse cells), fractions .25/.75.
// One thread per aligned fine pair (the two x-children of one coarse cell).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T lerpRef(T v0, T v1, T f) { return (T(1.) - f) * v0 + f * v1; }  // Ops.h:841-848

// A thread owns the two x-children of one coarse cell in TWO consecutive z-planes (the two z-children): they interpolate
// from the same 3 x 2 x 3 coarse values, which are loaded once.
template <typename T>
__device__ __forceinline__ void prolongBody(const TransferArgsT<T> &a, int vb, int tid)
{
    const int c = a.chunks[vb];
    const int zb = c / a.chunksPerPlane;
    const int64_t inPlane = int64_t(c - zb * a.chunksPerPlane) * CHUNK_CELLS + 2 * tid;
    // prologue: the fine labels of the thread's cells (static)
    uchar2 lab[CHUNK_Z];
#pragma unroll
    for (int q = 0; q < CHUNK_Z; ++q)
    {
	const int fz = zb * CHUNK_Z + q;
	lab[q] = make_uchar2(L_EXTERIOR, L_EXTERIOR);
	if (inPlane < a.finePlane && fz >= a.zlo && fz < a.zhi) lab[q] = *reinterpret_cast<const uchar2 *>(a.fineLabels + int64_t(fz) * a.finePlane + inPlane);
    }
    pdlWait();
    if (inPlane >= a.finePlane) return;
    const int fy = int(unsigned(inPlane) / unsigned(a.finePitch)), fx = int(inPlane - int64_t(fy) * a.finePitch);
    const int mx = (fx >> 1) + a.shift[0];
    const int my = (fy >> 1) + a.shift[1];
    const int ys = (fy & 1) ? my : my - 1;
    const T wy = (fy & 1) ? T(.25) : T(.75);
#pragma unroll
    for (int p = 0; p < CHUNK_Z / 2; ++p)
    {
	const int fz0 = zb * CHUNK_Z + 2 * p;  // even: storage origins are even
	const int64_t i0 = int64_t(fz0) * a.finePlane + inPlane, i1 = i0 + a.finePlane;
	const uchar2 l0 = lab[2 * p], l1 = lab[2 * p + 1];
	const bool a00 = (l0.x == L_INTERIOR || l0.x == L_BOUNDARY), a01 = (l0.y == L_INTERIOR || l0.y == L_BOUNDARY);
	const bool a10 = (l1.x == L_INTERIOR || l1.x == L_BOUNDARY), a11 = (l1.y == L_INTERIOR || l1.y == L_BOUNDARY);
	if (!(a00 | a01 | a10 | a11)) continue;
	const int mz = (fz0 >> 1) + a.shift[2];
	// x-lerps right after each row load (even x-child: start = m-1, f = .75; odd: start = m, f = .25), then y, then z:
	// the nesting and operation order of Ops.h:841-871, with 12 live values instead of 18
	T ex[2][3], ox[2][3];  // [y: ys, ys+1][z: mz-1..mz+1]
#pragma unroll
	for (int z = 0; z < 3; ++z)
#pragma unroll
	    for (int y = 0; y < 2; ++y)
	    {
		const T *row = a.coarse + (int64_t(mz - 1 + z) * a.coarsePlane + int64_t(ys + y) * a.coarsePitch + mx);
		const T v0 = row[-1], v1 = row[0], v2 = row[1];
		ex[y][z] = lerpRef<T>(v0, v1, T(.75));
		ox[y][z] = lerpRef<T>(v1, v2, T(.25));
	    }
	T ey[3], oy[3];
#pragma unroll
	for (int z = 0; z < 3; ++z)
	{
	    ey[z] = lerpRef<T>(ex[0][z], ex[1][z], wy);
	    oy[z] = lerpRef<T>(ox[0][z], ox[1][z], wy);
	}
	// the two z-children: even plane start = m-1, f = .75; odd plane start = m, f = .25
#pragma unroll
	for (int q = 0; q < 2; ++q)
	{
	    const bool ax = q ? a10 : a00, ay = q ? a11 : a01;
	    if (!(ax | ay)) continue;
	    const int64_t i = q ? i1 : i0;
	    const T wz = q ? T(.25) : T(.75);
	    const typename Vec2<T>::type old = ld2(a.out + i);
	    const T e = lerpRef<T>(ey[q], ey[q + 1], wz);
	    const T o = lerpRef<T>(oy[q], oy[q + 1], wz);
	    const T n0 = old.x + T(4.) * e, n1 = old.y + T(4.) * o;
	    if (ax & ay) st2(a.out + i, make2<T>(n0, n1));
	    else if (ax) a.out[i] = n0;
	    else a.out[i + 1] = n1;
	}
    }
}
template <typename T = double>
__global__ void __launch_bounds__(BLOCK, 6) k_prolong(const TransferArgsT<T> a) { pdlLaunch(); prolongBody<T>(a, blockIdx.x, threadIdx.x); }


// ------------------------------------------------------------------------------------------------
// Prolongation with the coarse values staged through shared memory by TMA: a 64 x 8 x 4 brick of fine cells interpolates
// from a 34 x 6 x 4 box of coarse cells (one bulk tensor copy, 6.5 KB), so the 18 scalar coarse loads per thread and z-pair of
// k_prolong -- each coarse value wanted by 3 x 2 threads -- come from shared memory and only the fine read-modify-write (the
// actual traffic: 17 of 18 bytes per cell) goes through the load/store unit to global memory.  Same arithmetic as prolongBody.
// ------------------------------------------------------------------------------------------------
constexpr int PB_BOX_X = TB_X / 2 + 2, PB_BOX_Y = TB_Y / 2 + 2, PB_BOX_Z = TB_Z / 2 + 2;
constexpr unsigned PB_BOX_BYTES = PB_BOX_X * PB_BOX_Y * PB_BOX_Z * sizeof(double);

__global__ void __launch_bounds__(BLOCK) k_prolong_tma(const TransferArgs a, const __grid_constant__ TmaMap tmCoarse, const int32_t *bricks, int bricksX, int bricksY, int ny)
{
    pdlLaunch();
    __shared__ alignas(128) double tile[PB_BOX_Z][PB_BOX_Y][PB_BOX_X];
    __shared__ alignas(8) unsigned long long bar;
    const int br = bricks[blockIdx.x];
    const int bz = br / (bricksX * bricksY), by = (br - bz * bricksX * bricksY) / bricksX, bx = br - (bz * bricksY + by) * bricksX;
    const int x0 = bx * TB_X, y0 = by * TB_Y, z0 = bz * TB_Z;
    const int px = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int fx = x0 + 2 * px, fy = y0 + ty;
    const bool inRow = fx < a.finePitch && fy < ny;
    const int64_t inPlane = int64_t(fy) * a.finePitch + fx;
    // prologue: the fine labels of the thread's cells (static)
    uchar2 lab[TB_Z];
#pragma unroll
    for (int q = 0; q < TB_Z; ++q)
    {
	const int fz = z0 + q;
	lab[q] = make_uchar2(L_EXTERIOR, L_EXTERIOR);
	if (inRow && fz >= a.zlo && fz < a.zhi) lab[q] = *reinterpret_cast<const uchar2 *>(a.fineLabels + int64_t(fz) * a.finePlane + inPlane);
    }
    if (threadIdx.x == 0) mbarInit(&bar, 1);
    __syncthreads();
    pdlWait();
    if (threadIdx.x == 0)
    {
	mbarExpectTx(&bar, PB_BOX_BYTES);
	// coarse cell of the brick's first fine cell, minus one in every direction; z + 1 = the coarse grid's guard plane
	tmaLoad3d(&tile[0][0][0], &tmCoarse, (x0 >> 1) + a.shift[0] - 1, (y0 >> 1) + a.shift[1] - 1, (z0 >> 1) + a.shift[2] - 1 + 1, &bar);
    }
    // the fine values to update: plain coalesced loads, in flight while the coarse box lands
    double2 old[TB_Z];
#pragma unroll
    for (int q = 0; q < TB_Z; ++q)
    {
	old[q] = make_double2(0.0, 0.0);
	const uchar2 l = lab[q];
	const bool any = (l.x == L_INTERIOR) | (l.x == L_BOUNDARY) | (l.y == L_INTERIOR) | (l.y == L_BOUNDARY);
	if (any) old[q] = ld2(a.out + int64_t(z0 + q) * a.finePlane + inPlane);
    }
    mbarWait(&bar, 0);
    const int ly = (ty & 1) ? (ty >> 1) + 1 : (ty >> 1);  // ys - box origin
    const double wy = (ty & 1) ? .25 : .75;               // fy and ty have the same parity (y0 is even)
#pragma unroll
    for (int p = 0; p < TB_Z / 2; ++p)
    {
	const uchar2 l0 = lab[2 * p], l1 = lab[2 * p + 1];
	const bool a00 = (l0.x == L_INTERIOR || l0.x == L_BOUNDARY), a01 = (l0.y == L_INTERIOR || l0.y == L_BOUNDARY);
	const bool a10 = (l1.x == L_INTERIOR || l1.x == L_BOUNDARY), a11 = (l1.y == L_INTERIOR || l1.y == L_BOUNDARY);
	if (!(a00 | a01 | a10 | a11)) continue;
	double ex[2][3], ox[2][3];  // [y: ys, ys+1][z: mz-1..mz+1]
#pragma unroll
	for (int z = 0; z < 3; ++z)
#pragma unroll
	    for (int y = 0; y < 2; ++y)
	    {
		const double *row = &tile[p + z][ly + y][px + 1];
		const double v0 = row[-1], v1 = row[0], v2 = row[1];
		ex[y][z] = lerpRef(v0, v1, .75);
		ox[y][z] = lerpRef(v1, v2, .25);
	    }
	double ey[3], oy[3];
#pragma unroll
	for (int z = 0; z < 3; ++z)
	{
	    ey[z] = lerpRef(ex[0][z], ex[1][z], wy);
	    oy[z] = lerpRef(ox[0][z], ox[1][z], wy);
	}
#pragma unroll
	for (int q = 0; q < 2; ++q)
	{
	    const bool ax = q ? a10 : a00, ay = q ? a11 : a01;
	    if (!(ax | ay)) continue;
	    const int64_t i = int64_t(z0 + 2 * p + q) * a.finePlane + inPlane;
	    const double wz = q ? .25 : .75;
	    const double2 o = old[2 * p + q];
	    const double e = lerpRef(ey[q], ey[q + 1], wz);
	    const double od = lerpRef(oy[q], oy[q + 1], wz);
	    const double n0 = o.x + 4. * e, n1 = o.y + 4. * od;
	    if (ax & ay) st2(a.out + i, make_double2(n0, n1));
	    else if (ax) a.out[i] = n0;
	    else a.out[i + 1] = n1;
	}
    }
}

// brick flags for the prolongation: 1 when the 64 x 8 x 4 brick holds an ACTIVE cell
__global__ void __launch_bounds__(BLOCK) k_brick_flags_active(uint8_t *flags, const uint8_t *labels, int bricksX, int bricksY, int pitch, int64_t plane, int ny, int nz)
{
    const int br = blockIdx.x;
    const int bz = br / (bricksX * bricksY), by = (br - bz * bricksX * bricksY) / bricksX, bx = br - (bz * bricksY + by) * bricksX;
    const int x = bx * TB_X + 2 * (threadIdx.x & 31), y = by * TB_Y + (threadIdx.x >> 5);
    int any = 0;
    if (x < pitch && y < ny)
	for (int dz = 0; dz < TB_Z; ++dz)
	{
	    const int z = bz * TB_Z + dz;
	    if (z >= nz) break;
	    const uchar2 l = *reinterpret_cast<const uchar2 *>(labels + int64_t(z) * plane + int64_t(y) * pitch + x);
	    any |= (l.x == L_INTERIOR) | (l.x == L_BOUNDARY) | (l.y == L_INTERIOR) | (l.y == L_BOUNDARY);
	}
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) flags[br] = uint8_t(any);
}

// ------------------------------------------------------------------------------------------------
// Coarsest level (MG.cpp:669-692): gather b, x = A^-1 b (dense inverse of the SPD matrix, built on the
// host at setup from an exact Cholesky factor), scatter.  One warp per row, b staged in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_coarse_solve(double *x, const double *b, const int32_t *idx, const double *inv, int n)
{ pdlEnter();
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
    int nChunks;        // CTAs beyond it do no cell work (a reduction on a slab without active cells still launches ONE CTA, so that
			// its result -- 0 -- and the scalars the finishing CTA maintains are written on every rank)
    int chunksPerPlane;
    int64_t plane;
    int nz;
    int zlo, zhi;       // z-plane clip [zlo, zhi)
    int redLo, redHi;   // planes that enter a fused reduction (the rank's owned planes)
    double *y;          // destination / first operand
    const double *a;    // second operand
    const double *c;    // third operand
    double *y2;         // second destination (fused CG update)
    double s;           // host scalar
    Scalars *sc;        // device scalars
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
    VO_CG_UPDATE,      // alpha = rhoNew/pAp; y(x) += alpha*a(p); y2(r) -= alpha*c(Ap); result = |r|^2; then rho = rhoNew
    VO_CG_DIRECTION,   // beta = rhoNew/rho; y(p) = a(z) + beta*y(p)
    VO_COPY,           // y = a
    VO_MUL             // y = a * c  (diagonal preconditioner, GFS.cpp:562-603)
};

template <int OP>
__global__ void __launch_bounds__(BLOCK) k_vec(const VecArgs v)
{
    pdlLaunch();
    const bool live = int(blockIdx.x) < v.nChunks;
    const int c = live ? v.chunks[blockIdx.x] : 0;
    pdlWait();
    const int zb = c / v.chunksPerPlane;
    const int64_t inPlane = int64_t(c - zb * v.chunksPerPlane) * CHUNK_CELLS + 2 * threadIdx.x;
    double acc = 0.0;
    double s = v.s;
    if (OP == VO_CG_UPDATE) s = v.sc->rhoNew / v.sc->pAp;
    if (OP == VO_CG_DIRECTION) s = v.sc->rhoNew / v.sc->rho;
    if (live && inPlane < v.plane)
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
		if (z >= v.redLo && z < v.redHi) { acc += r.x * r.x; acc += r.y * r.y; }
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
	    else if (OP == VO_MUL)
	    {
		const double2 a = ld2(v.a + i), cc = ld2(v.c + i);
		st2(v.y + i, make_double2(a.x * cc.x, a.y * cc.y));
	    }
	}
    }
    if (OP == VO_DOT || OP == VO_NORM2) gridReduce<false>(acc, v.partials, v.ticket, v.result);
    if (OP == VO_CG_UPDATE)
    {
	// the CTA that finishes the reduction also retires this iteration's rho (CG.h:165-176: absOld = absNew): every CTA has
	// read rhoNew by then, and nothing reads rho before the next direction update
	if (gridReduce<false>(acc, v.partials, v.ticket, v.result)) v.sc->rho = v.sc->rhoNew;
    }
    if (OP == VO_MAX) gridReduce<true>(acc, v.partials, v.ticket, v.result);
}

// fp64 <-> fp32 over the active chunks (mixed precision: the V-cycle's input and output cross the precision boundary once each)
template <typename TO, typename FROM>
__global__ void __launch_bounds__(BLOCK) k_convert(TO *dst, const FROM *src, const int32_t *chunks, int chunksPerPlane, int64_t plane, int nz)
{
    pdlLaunch();
    const int c = chunks[blockIdx.x];
    pdlWait();
    const int zb = c / chunksPerPlane;
    const int64_t inPlane = int64_t(c - zb * chunksPerPlane) * CHUNK_CELLS + 2 * threadIdx.x;
    if (inPlane >= plane) return;
#pragma unroll
    for (int dz = 0; dz < CHUNK_Z; ++dz)
    {
	const int z = zb * CHUNK_Z + dz;
	if (z >= nz) break;
	const int64_t i = int64_t(z) * plane + inPlane;
	const typename Vec2<FROM>::type v = ld2(src + i);
	st2(dst + i, make2<TO>(TO(v.x), TO(v.y)));
    }
}

// zero the active chunks of a grid (x = 0 at the start of a V-cycle level, MG.cpp:439-440, :566)
__device__ __forceinline__ void zeroBody(double *y, int c, int chunksPerPlane, int64_t plane, int nz, int tid)
{
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
    pdlLaunch();
    const int c = chunks[blockIdx.x];
    pdlWait();
    zeroBody(y, c, chunksPerPlane, plane, nz, threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// Tiled Gauss-Seidel (Ops.h:369-520), the reference's production smoother (GFS.cpp:463-466).  The reference sweeps the
// 16^3 tiles of one parity of (tx+ty+tz) in parallel and the cells of a tile one after the other in lexicographic order
// (x fastest), forwards or backwards, undamped: u += (b - A u) / diag.  Face-adjacent tiles have opposite parity, so a
// tile only sees frozen values outside itself; inside, cell (x,y,z) depends on the NEW values of (x-1,y,z), (x,y-1,z),
// (x,y,z-1) and the OLD values of the +1 neighbours (mirrored for the backward sweep).  All cells of one wavefront
// x+y+z = s are therefore independent, and sweeping s = 0..45 (or 45..0) reproduces the sequential result exactly.
// One CTA per tile: the tile and its one-cell halo live in shared memory for the 46 steps.
// ------------------------------------------------------------------------------------------------
constexpr int GS_TILE = 16;
constexpr int GS_HALO = GS_TILE + 2;

struct GsArgs
{
    double *x;
    const double *b;
    const uint8_t *labels;
    const int32_t *tiles;    // linear tile ids (x fastest) of the active tiles of the wanted parity
    const int32_t *bpos;     // grid: boundary-record index of a BOUNDARY cell (undefined elsewhere)
    const double *bcoef;
    int nBoundary;
    int tilesX, tilesY;      // tile grid over the storage box
    int off[3];              // storage coordinate of tile (0,0,0)'s first cell (<= 0)
    int n[3];
    int pitch;
    int64_t plane;
    int forward;
};

__global__ void __launch_bounds__(BLOCK) k_gauss_seidel(const GsArgs a)
{ pdlEnter();
    extern __shared__ double gsm[];
    double *xs = gsm;                                   // [18][18][18]
    double *bs = gsm + GS_HALO * GS_HALO * GS_HALO;     // [16][16][16]
    uint8_t *ls = reinterpret_cast<uint8_t *>(bs + GS_TILE * GS_TILE * GS_TILE);  // [16][16][16]
    const int t = a.tiles[blockIdx.x];
    const int tz = t / (a.tilesX * a.tilesY), ty = (t - tz * a.tilesX * a.tilesY) / a.tilesX, tx = t - (tz * a.tilesY + ty) * a.tilesX;
    const int ox = a.off[0] + tx * GS_TILE, oy = a.off[1] + ty * GS_TILE, oz = a.off[2] + tz * GS_TILE;
    for (int i = threadIdx.x; i < GS_HALO * GS_HALO * GS_HALO; i += BLOCK)
    {
	const int lz = i / (GS_HALO * GS_HALO), ly = (i - lz * GS_HALO * GS_HALO) / GS_HALO, lx = i - (lz * GS_HALO + ly) * GS_HALO;
	const int gx = ox + lx - 1, gy = oy + ly - 1, gz = oz + lz - 1;
	double v = 0.0;
	if (gx >= 0 && gy >= 0 && gz >= 0 && gx < a.n[0] && gy < a.n[1] && gz < a.n[2]) v = a.x[int64_t(gz) * a.plane + int64_t(gy) * a.pitch + gx];
	xs[i] = v;
    }
    for (int i = threadIdx.x; i < GS_TILE * GS_TILE * GS_TILE; i += BLOCK)
    {
	const int lz = i >> 8, ly = (i >> 4) & 15, lx = i & 15;
	const int gx = ox + lx, gy = oy + ly, gz = oz + lz;
	double v = 0.0;
	uint8_t l = L_EXTERIOR;
	if (gx >= 0 && gy >= 0 && gz >= 0 && gx < a.n[0] && gy < a.n[1] && gz < a.n[2])
	{
	    const int64_t g = int64_t(gz) * a.plane + int64_t(gy) * a.pitch + gx;
	    l = a.labels[g];
	    if (l == L_INTERIOR || l == L_BOUNDARY) v = a.b[g];
	}
	bs[i] = v;
	ls[i] = l;
    }
    __syncthreads();
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    for (int step = 0; step <= 3 * (GS_TILE - 1); ++step)
    {
	const int sfront = a.forward ? step : 3 * (GS_TILE - 1) - step;
	const int lz = sfront - lx - ly;
	if (lz >= 0 && lz < GS_TILE)
	{
	    const int li = (lz << 8) | (ly << 4) | lx;
	    const int l = ls[li];
	    if (l == L_INTERIOR || l == L_BOUNDARY)
	    {
		const int c = ((lz + 1) * GS_HALO + (ly + 1)) * GS_HALO + (lx + 1);
		const double u[6] = {xs[c - 1], xs[c + 1], xs[c - GS_HALO], xs[c + GS_HALO], xs[c - GS_HALO * GS_HALO], xs[c + GS_HALO * GS_HALO]};
		const double centre = xs[c];
		double lap = 0.0, diag = 6.0;
		if (l == L_INTERIOR)
		{
#pragma unroll
		    for (int n = 0; n < 6; ++n) lap -= u[n];
		}
		else
		{
		    const int64_t g = int64_t(oz + lz) * a.plane + int64_t(oy + ly) * a.pitch + (ox + lx);
		    const int k = a.bpos[g];
#pragma unroll
		    for (int n = 0; n < 6; ++n)
		    {
			const double cn = a.bcoef[int64_t(n) * a.nBoundary + k];
			if (cn != 0.0) lap -= cn * u[n];
		    }
		    diag = a.bcoef[int64_t(6) * a.nBoundary + k];
		}
		lap += diag * centre;
		double r = bs[li] - lap;   // Ops.h:490-493
		r /= diag;
		xs[c] = centre + r;
	    }
	}
	__syncthreads();
    }
    for (int i = threadIdx.x; i < GS_TILE * GS_TILE * GS_TILE; i += BLOCK)
    {
	const int l = ls[i];
	if (!(l == L_INTERIOR || l == L_BOUNDARY)) continue;
	const int lz = i >> 8, ly2 = (i >> 4) & 15, lx2 = i & 15;
	a.x[int64_t(oz + lz) * a.plane + int64_t(oy + ly2) * a.pitch + (ox + lx2)] = xs[((lz + 1) * GS_HALO + (ly2 + 1)) * GS_HALO + (lx2 + 1)];
    }
}


// ------------------------------------------------------------------------------------------------
// The same half-pass with nothing but shared memory on the 46-step critical path.  k_gauss_seidel above fetches a BOUNDARY
// cell's record inside the wavefront step that updates it -- bpos[cell] -> coefficient rows, two dependent global round trips of
// ~0.7 us each, paid by the whole CTA at the step's barrier -- so a tile on the free surface costs 46 x a few microseconds.
// Here every record a tile needs is gathered BEFORE the sweep: each BOUNDARY cell of the tile takes a slot in a shared-memory
// pool holding its diagonal and its 2-bit coefficient codes (k_band_coef: coefficient 0 / exactly 1 / fractional); a step is
// then shared-memory loads, fp64 arithmetic and one barrier.  Fractional coefficients (rare: see coefCode) and pool overflow
// still go to global memory.  Same arithmetic, same order: bitwise k_gauss_seidel.
// ------------------------------------------------------------------------------------------------
constexpr int GS_POOL = 1280;               // BOUNDARY records per tile held in shared memory (a 16 x 16 surface patch 3 cells thick is 768)
constexpr unsigned short GS_SLOT_NONE = 0xffffu, GS_SLOT_OVERFLOW = 0xfffeu;

struct GsArgs2
{
    GsArgs g;
    const unsigned short *wcode;
};

__global__ void __launch_bounds__(BLOCK, 2) k_gauss_seidel2(const GsArgs2 p)
{
    pdlLaunch();
    const GsArgs &a = p.g;
    extern __shared__ double gsm[];
    double *xs = gsm;                                                     // [18][18][18]
    double *bs = xs + GS_HALO * GS_HALO * GS_HALO;                        // [16][16][16]
    double *poolDiag = bs + GS_TILE * GS_TILE * GS_TILE;                  // [GS_POOL]
    unsigned short *poolCode = reinterpret_cast<unsigned short *>(poolDiag + GS_POOL);  // [GS_POOL]
    unsigned short *slot = poolCode + GS_POOL;                            // [16][16][16] pool slot of a BOUNDARY cell
    uint8_t *ls = reinterpret_cast<uint8_t *>(slot + GS_TILE * GS_TILE * GS_TILE);      // [16][16][16]
    __shared__ int poolCount, frontLo, frontHi;
    const int t = a.tiles[blockIdx.x];
    const int tz = t / (a.tilesX * a.tilesY), ty = (t - tz * a.tilesX * a.tilesY) / a.tilesX, tx = t - (tz * a.tilesY + ty) * a.tilesX;
    const int ox = a.off[0] + tx * GS_TILE, oy = a.off[1] + ty * GS_TILE, oz = a.off[2] + tz * GS_TILE;
    if (threadIdx.x == 0) { poolCount = 0; frontLo = 3 * GS_TILE; frontHi = -1; }
    __syncthreads();
    // A thread owns the column (lx, ly) of the tile: cell i = q * 256 + tid is (lx, ly, lz = q).  Every phase below issues the
    // column's 16 loads TOGETHER (unrolled, predicated) -- the first version walked them one dependent round trip at a time
    // (label -> position -> record, 16 times over) and a tile spent half of its 60 us there (profiles/r02_gauss_seidel.md).
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const int gx0 = ox + lx, gy0 = oy + ly;
    const bool inXY = gx0 >= 0 && gy0 >= 0 && gx0 < a.n[0] && gy0 < a.n[1];
    const int64_t g0 = int64_t(oz) * a.plane + int64_t(gy0) * a.pitch + gx0;
    // prologue (static): labels, and the records of the tile's BOUNDARY cells into the pool
    uint8_t lab[GS_TILE];
#pragma unroll
    for (int q = 0; q < GS_TILE; ++q)
    {
	const int gz = oz + q;
	lab[q] = (inXY && gz >= 0 && gz < a.n[2]) ? a.labels[g0 + int64_t(q) * a.plane] : uint8_t(L_EXTERIOR);
    }
    int bp[GS_TILE];
#pragma unroll
    for (int q = 0; q < GS_TILE; ++q) bp[q] = lab[q] == L_BOUNDARY ? a.bpos[g0 + int64_t(q) * a.plane] : -1;
    {
	// the wavefronts lx + ly + lz = s that hold an active cell of this tile: a front without one changes nothing, so the sweep
	// below only walks [frontLo, frontHi] (a tile of a coarse level is mostly EXTERIOR: 22 of the 46 fronts at 8^3 active cells)
	int lo = 3 * GS_TILE, hi = -1;
#pragma unroll
	for (int q = 0; q < GS_TILE; ++q)
	    if (lab[q] == L_INTERIOR || lab[q] == L_BOUNDARY)
	    {
		lo = min(lo, lx + ly + q);
		hi = max(hi, lx + ly + q);
	    }
	lo = __reduce_min_sync(0xffffffffu, lo);
	hi = __reduce_max_sync(0xffffffffu, hi);
	if ((threadIdx.x & 31) == 0 && hi >= 0)
	{
	    atomicMin(&frontLo, lo);
	    atomicMax(&frontHi, hi);
	}
    }
    double dg[GS_TILE];
    unsigned short cd[GS_TILE];
#pragma unroll
    for (int q = 0; q < GS_TILE; ++q)
    {
	dg[q] = bp[q] >= 0 ? a.bcoef[int64_t(6) * a.nBoundary + bp[q]] : 0.0;
	cd[q] = bp[q] >= 0 ? p.wcode[bp[q]] : (unsigned short)0;
    }
#pragma unroll
    for (int q = 0; q < GS_TILE; ++q)
    {
	unsigned short sl = GS_SLOT_NONE;
	if (bp[q] >= 0)
	{
	    const int s0 = atomicAdd(&poolCount, 1);
	    if (s0 < GS_POOL)
	    {
		poolDiag[s0] = dg[q];
		poolCode[s0] = cd[q];
		sl = (unsigned short)s0;
	    }
	    else sl = GS_SLOT_OVERFLOW;
	}
	ls[q * BLOCK + threadIdx.x] = lab[q];
	slot[q * BLOCK + threadIdx.x] = sl;
    }
    pdlWait();
    {
	constexpr int HALO_CELLS = GS_HALO * GS_HALO * GS_HALO, HALO_ROUNDS = (HALO_CELLS + BLOCK - 1) / BLOCK;
	double hv[HALO_ROUNDS];
#pragma unroll
	for (int q = 0; q < HALO_ROUNDS; ++q)
	{
	    const int i = q * BLOCK + int(threadIdx.x);
	    const int hz = i / (GS_HALO * GS_HALO), hy = (i - hz * GS_HALO * GS_HALO) / GS_HALO, hx = i - (hz * GS_HALO + hy) * GS_HALO;
	    const int gx = ox + hx - 1, gy = oy + hy - 1, gz = oz + hz - 1;
	    hv[q] = 0.0;
	    if (i < HALO_CELLS && gx >= 0 && gy >= 0 && gz >= 0 && gx < a.n[0] && gy < a.n[1] && gz < a.n[2])
		hv[q] = a.x[int64_t(gz) * a.plane + int64_t(gy) * a.pitch + gx];
	}
	double bv[GS_TILE];
#pragma unroll
	for (int q = 0; q < GS_TILE; ++q)
	    bv[q] = (lab[q] == L_INTERIOR || lab[q] == L_BOUNDARY) ? a.b[g0 + int64_t(q) * a.plane] : 0.0;
#pragma unroll
	for (int q = 0; q < HALO_ROUNDS; ++q)
	{
	    const int i = q * BLOCK + int(threadIdx.x);
	    if (i < HALO_CELLS) xs[i] = hv[q];
	}
#pragma unroll
	for (int q = 0; q < GS_TILE; ++q) bs[q * BLOCK + threadIdx.x] = bv[q];
    }
    __syncthreads();
    const int firstFront = frontLo, lastFront = frontHi;
    for (int step = 0; step <= lastFront - firstFront; ++step)
    {
	const int sfront = a.forward ? firstFront + step : lastFront - step;
	const int lz = sfront - lx - ly;
	if (lz >= 0 && lz < GS_TILE)
	{
	    const int li = (lz << 8) | (ly << 4) | lx;
	    const int l = ls[li];
	    if (l == L_INTERIOR || l == L_BOUNDARY)
	    {
		const int c = ((lz + 1) * GS_HALO + (ly + 1)) * GS_HALO + (lx + 1);
		const double u[6] = {xs[c - 1], xs[c + 1], xs[c - GS_HALO], xs[c + GS_HALO], xs[c - GS_HALO * GS_HALO], xs[c + GS_HALO * GS_HALO]};
		const double centre = xs[c];
		double lap = 0.0, diag = 6.0;
		if (l == L_INTERIOR)
		{
#pragma unroll
		    for (int n = 0; n < 6; ++n) lap -= u[n];
		}
		else
		{
		    const unsigned sl = slot[li];
		    if (sl < unsigned(GS_POOL))
		    {
			const unsigned code = poolCode[sl];
			diag = poolDiag[sl];
#pragma unroll
			for (int n = 0; n < 6; ++n)
			{
			    const unsigned cc = (code >> (2 * n)) & 3u;
			    if (cc == 1u) lap -= u[n];
			    else if (cc == 2u)
			    {
				const int64_t g = int64_t(oz + lz) * a.plane + int64_t(oy + ly) * a.pitch + (ox + lx);
				lap -= a.bcoef[int64_t(n) * a.nBoundary + a.bpos[g]] * u[n];
			    }
			}
		    }
		    else
		    {
			const int64_t g = int64_t(oz + lz) * a.plane + int64_t(oy + ly) * a.pitch + (ox + lx);
			const int k = a.bpos[g];
#pragma unroll
			for (int n = 0; n < 6; ++n)
			{
			    const double cn = a.bcoef[int64_t(n) * a.nBoundary + k];
			    if (cn != 0.0) lap -= cn * u[n];
			}
			diag = a.bcoef[int64_t(6) * a.nBoundary + k];
		    }
		}
		lap += diag * centre;
		double r = bs[li] - lap;   // Ops.h:490-493
		r /= diag;
		xs[c] = centre + r;
	    }
	}
	__syncthreads();
    }
    for (int i = threadIdx.x; i < GS_TILE * GS_TILE * GS_TILE; i += BLOCK)
    {
	const int l = ls[i];
	if (!(l == L_INTERIOR || l == L_BOUNDARY)) continue;
	const int lz = i >> 8, ly2 = (i >> 4) & 15, lx2 = i & 15;
	a.x[int64_t(oz + lz) * a.plane + int64_t(oy + ly2) * a.pitch + (ox + lx2)] = xs[((lz + 1) * GS_HALO + (ly2 + 1)) * GS_HALO + (lx2 + 1)];
    }
}

// tile flags for the two parities: bit set when the tile holds an active cell
__global__ void __launch_bounds__(BLOCK) k_gs_tile_flags(uint8_t *flagOdd, uint8_t *flagEven, const uint8_t *labels, int tilesX, int tilesY, int off0,
							int off1, int off2, int n0, int n1, int n2, int pitch, int64_t plane, int parity0)
{
    const int t = blockIdx.x;
    const int tz = t / (tilesX * tilesY), ty = (t - tz * tilesX * tilesY) / tilesX, tx = t - (tz * tilesY + ty) * tilesX;
    const int ox = off0 + tx * GS_TILE, oy = off1 + ty * GS_TILE, oz = off2 + tz * GS_TILE;
    int any = 0;
    for (int i = threadIdx.x; i < GS_TILE * GS_TILE * GS_TILE; i += BLOCK)
    {
	const int gx = ox + (i & 15), gy = oy + ((i >> 4) & 15), gz = oz + (i >> 8);
	if (gx >= 0 && gy >= 0 && gz >= 0 && gx < n0 && gy < n1 && gz < n2)
	{
	    const int l = labels[int64_t(gz) * plane + int64_t(gy) * pitch + gx];
	    any |= (l == L_INTERIOR || l == L_BOUNDARY);
	}
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0)
    {
	const int odd = (parity0 + tx + ty + tz) & 1;
	flagOdd[t] = uint8_t(any && odd);
	flagEven[t] = uint8_t(any && !odd);
    }
}

// ------------------------------------------------------------------------------------------------
// Compact coarse sub-V-cycle.  Below a few thousand cells a level is pure latency: 19 dependent launches of ~3 us, each
// a chain of L2 round trips.  Levels [first, last] -- down-stroke, direct solve, up-stroke -- therefore run in ONE CTA
// whose vectors (x, rhs, scratch of every level) AND index tables live in SHARED MEMORY: compact lists of the active
// cells with precomputed 16-bit neighbour / restriction / prolongation indices, copied in once per launch.  A step
// costs a __syncthreads instead of a kernel boundary and touches no global memory.  Per cell the arithmetic and its order are those of k_stencil / k_band /
// k_restrict / k_prolong / k_coarse_solve, so the result is bitwise identical to the per-kernel path.
// ------------------------------------------------------------------------------------------------
constexpr int CYCLE_THREADS = 1024;
constexpr int CYCLE_MAX_LEVELS = 8;
constexpr unsigned short CYCLE_NONE = 0xffffu;

// Byte offsets into the table blob (copied into shared memory at kernel start; 16-bit compact indices, CYCLE_NONE = not active)
struct CompactLevel
{
    int n;         // active cells of the level (< 65535)
    int off;       // offset (in doubles) of the level's x | t | b arrays in shared memory
    int nbr;       // u16 [6][n]  compact index of the -x,+x,-y,+y,-z,+z neighbour
    int rst;       // u16 [64][n] compact indices (in the next FINER level) of the 4x4x4 restriction taps
    int pro;       // u16 [8][n]  compact indices (in the next COARSER level) of the 2x2x2 prolongation corners
    int diag;      // u8  [n]     6 for INTERIOR; number of non-EXTERIOR neighbours for BOUNDARY (Ops.h:237-248, weight 1)
    int flags;     // u8  [n]     bit0 = cell belongs to the boundary band, bits 1..3 = parity of its x, y, z storage index
};

struct CompactArgs
{
    CompactLevel lv[CYCLE_MAX_LEVELS];  // lv[0] is the finest compact level
    int nLevels;
    int sweeps;
    int vectorDoubles;        // doubles of the vector region; the table blob follows it in shared memory
    int blobBytes;            // multiple of 16
    int solveIdx;             // u16 [nSolve] compact index of the direct solve's k-th unknown (MG.cpp:296-323 numbering)
    int nSolve;
    const unsigned char *blob;
    const int32_t *cellTop;   // [lv[0].n] storage index of lv[0]'s cells in its grid
    const double *bTop;       // rhs grid of lv[0] (written by the restriction kernel of the level above)
    double *xTop;             // solution grid of lv[0] (read by the prolongation kernel of the level above)
    const float *bTop32;      // mixed precision: the same two grids in fp32 (then bTop / xTop are unused); the cycle itself stays fp64
    float *xTop32;
    const double *inv;        // [nSolve][nSolve]
};

// A x at compact cell k, in the operation order of computeLaplacian (Ops.h:177-260): neighbours by (axis, direction), centre last
__device__ __forceinline__ double compactLap(const unsigned short *nbr, int n, const double *x, int k, double diag)
{
    double lap = 0.0;
#pragma unroll
    for (int d = 0; d < 6; ++d)
    {
	const unsigned short j = nbr[d * n + k];
	if (j != CYCLE_NONE) lap -= x[j];
    }
    lap += diag * x[k];
    return lap;
}

// one damped-Jacobi sweep over the band cells (bandOnly) or all cells: t = new values, then x = t (Ops.h:262-367, :524-619)
__device__ __forceinline__ void compactSweep(const CompactLevel &L, const unsigned char *tab, double *x, double *t, const double *b, bool bandOnly)
{
    const unsigned short *nbr = reinterpret_cast<const unsigned short *>(tab + L.nbr);
    const unsigned char *dg = tab + L.diag, *fl = tab + L.flags;
    for (int k = threadIdx.x; k < L.n; k += CYCLE_THREADS)
    {
	if (bandOnly && !(fl[k] & 1)) continue;
	const double diag = double(dg[k]);
	const double lap = compactLap(nbr, L.n, x, k, diag);
	double r = b[k] - lap;
	r /= diag;
	t[k] = x[k] + (2.0 / 3.0) * r;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < L.n; k += CYCLE_THREADS)
    {
	if (bandOnly && !(fl[k] & 1)) continue;
	x[k] = t[k];
    }
    __syncthreads();
}

__device__ __forceinline__ void compactSmooth(const CompactLevel &L, const unsigned char *tab, double *x, double *t, const double *b, int sweeps)
{
    for (int s = 0; s < sweeps; ++s) compactSweep(L, tab, x, t, b, true);
    compactSweep(L, tab, x, t, b, false);
    for (int s = 0; s < sweeps; ++s) compactSweep(L, tab, x, t, b, true);
}

__global__ void __launch_bounds__(CYCLE_THREADS, 1) k_compact_cycle(const CompactArgs c)
{
    pdlLaunch();
    extern __shared__ double sm[];
    unsigned char *tab = reinterpret_cast<unsigned char *>(sm + c.vectorDoubles);
    const int nl = c.nLevels;
    {
	// prologue: the (static) index tables go to shared memory while the restriction above this level is still running
	const int4 *src = reinterpret_cast<const int4 *>(c.blob);
	int4 *dst = reinterpret_cast<int4 *>(tab);
	for (int i = threadIdx.x; i < c.blobBytes / 16; i += CYCLE_THREADS) dst[i] = __ldg(src + i);
	pdlWait();
	const CompactLevel &L = c.lv[0];
	double *b = sm + L.off + 2 * L.n;
	if (c.bTop32)
	    for (int k = threadIdx.x; k < L.n; k += CYCLE_THREADS) b[k] = double(c.bTop32[__ldg(c.cellTop + k)]);
	else
	    for (int k = threadIdx.x; k < L.n; k += CYCLE_THREADS) b[k] = c.bTop[__ldg(c.cellTop + k)];
    }
    // ---- down-stroke (MG.cpp:557-667): x = 0, smooth, residual, restrict
    for (int l = 0; l + 1 < nl; ++l)
    {
	const CompactLevel &L = c.lv[l];
	const CompactLevel &C = c.lv[l + 1];
	double *x = sm + L.off, *t = x + L.n, *b = t + L.n;
	for (int k = threadIdx.x; k < L.n; k += CYCLE_THREADS) x[k] = 0.0;
	__syncthreads();
	compactSmooth(L, tab, x, t, b, c.sweeps);
	{
	    const unsigned short *nbr = reinterpret_cast<const unsigned short *>(tab + L.nbr);
	    for (int k = threadIdx.x; k < L.n; k += CYCLE_THREADS)
	    {
		const double lap = compactLap(nbr, L.n, x, k, double(tab[L.diag + k]));
		t[k] = b[k] + (-1.0) * lap;  // Ops.h:731
	    }
	}
	__syncthreads();
	double *bc = sm + C.off + 2 * C.n;
	const unsigned short *rst = reinterpret_cast<const unsigned short *>(tab + C.rst);
	for (int k = threadIdx.x; k < C.n; k += CYCLE_THREADS)
	{
	    const double rw[4] = {1. / 8., 3. / 8., 3. / 8., 1. / 8.};
	    double v = 0.0;
#pragma unroll
	    for (int z = 0; z < 4; ++z)
#pragma unroll
		for (int y = 0; y < 4; ++y)
#pragma unroll
		    for (int xx = 0; xx < 4; ++xx)
		    {
			const unsigned short j = rst[((z * 4 + y) * 4 + xx) * C.n + k];
			if (j != CYCLE_NONE) v += rw[xx] * rw[y] * rw[z] * t[j];  // an inactive tap adds +0.0, which never changes v
		    }
	    bc[k] = v;
	}
	__syncthreads();
    }
    // ---- direct solve on the coarsest level (MG.cpp:669-692): x = A^-1 b, one warp per row as in k_coarse_solve
    {
	const CompactLevel &L = c.lv[nl - 1];
	double *x = sm + L.off, *t = x + L.n, *b = t + L.n;
	const unsigned short *sidx = reinterpret_cast<const unsigned short *>(tab + c.solveIdx);
	const int n = c.nSolve;
	for (int i = threadIdx.x; i < n; i += CYCLE_THREADS) t[i] = b[sidx[i]];
	__syncthreads();
	const int lane = threadIdx.x & 31;
	for (int row = threadIdx.x >> 5; row < n; row += CYCLE_THREADS / 32)
	{
	    const double *r = c.inv + int64_t(row) * n;
	    double acc = 0.0;
	    for (int j = lane; j < n; j += 32) acc += r[j] * t[j];
	    acc = warpSum(acc);
	    if (lane == 0) x[sidx[row]] = acc;
	}
	__syncthreads();
    }
    // ---- up-stroke (MG.cpp:695-784): x += 4 trilerp(x_coarse), smooth
    for (int l = nl - 2; l >= 0; --l)
    {
	const CompactLevel &L = c.lv[l];
	const CompactLevel &C = c.lv[l + 1];
	double *x = sm + L.off, *t = x + L.n, *b = t + L.n;
	const double *xc = sm + C.off;
	const unsigned short *pro = reinterpret_cast<const unsigned short *>(tab + L.pro);
	for (int k = threadIdx.x; k < L.n; k += CYCLE_THREADS)
	{
	    const int f = tab[L.flags + k];
	    const double wx = (f & 2) ? .25 : .75, wy = (f & 4) ? .25 : .75, wz = (f & 8) ? .25 : .75;
	    double v[8];
#pragma unroll
	    for (int q = 0; q < 8; ++q)
	    {
		const unsigned short j = pro[q * L.n + k];
		v[q] = j != CYCLE_NONE ? xc[j] : 0.0;
	    }
	    // corner q = x + 2 y + 4 z; lerp nesting x -> y -> z (Ops.h:841-871)
	    const double e = lerpRef<double>(lerpRef<double>(lerpRef<double>(v[0], v[1], wx), lerpRef<double>(v[2], v[3], wx), wy),
				     lerpRef<double>(lerpRef<double>(v[4], v[5], wx), lerpRef<double>(v[6], v[7], wx), wy), wz);
	    x[k] = x[k] + 4. * e;
	}
	__syncthreads();
	compactSmooth(L, tab, x, t, b, c.sweeps);
    }
    {
	const CompactLevel &L = c.lv[0];
	const double *x = sm + L.off;
	if (c.xTop32)
	    for (int k = threadIdx.x; k < L.n; k += CYCLE_THREADS) c.xTop32[__ldg(c.cellTop + k)] = float(x[k]);
	else
	    for (int k = threadIdx.x; k < L.n; k += CYCLE_THREADS) c.xTop[__ldg(c.cellTop + k)] = x[k];
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

// staging int32 / uint8 (dense box, x-fastest, row pitch = n[0]) -> byte labels with EXTERIOR outside [validLo, validHi)
template <typename LabelT>
__global__ void __launch_bounds__(BLOCK) k_labels_from_host(uint8_t *labels, const LabelT *staging, BoxArgs g, int lo0, int lo1, int lo2,
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
// band flags of SM_JACOBI_ZERO: bit0 = in the band, bit1 = in the band dilated once more (INTERIOR cells only)
__global__ void __launch_bounds__(BLOCK) k_band_flag_grid(uint8_t *flags, const uint8_t *band, const uint8_t *dilated, int64_t total)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i < total) flags[i] = uint8_t((band[i] ? 1 : 0) | (dilated[i] ? 2 : 0));
}
// neighbour mask of SM_JACOBI_ZERO: bit0 = the cell is in the band; on INTERIOR cells (never on the storage border) bits 1..6 = the
// -x, +x, -y, +y, -z, +z neighbour is in the band
__global__ void __launch_bounds__(BLOCK) k_band_nbr_mask(uint8_t *mask, const uint8_t *band, const uint8_t *labels, BoxArgs g)
{
    const int64_t i = int64_t(blockIdx.x) * BLOCK + threadIdx.x;
    if (i >= g.total) return;
    unsigned m = band[i] ? 1u : 0u;
    if (labels[i] == L_INTERIOR)
	m |= (band[i - 1] ? 2u : 0u) | (band[i + 1] ? 4u : 0u) | (band[i - g.pitch] ? 8u : 0u) | (band[i + g.pitch] ? 16u : 0u) | (band[i - g.plane] ? 32u : 0u) |
	     (band[i + g.plane] ? 64u : 0u);
    mask[i] = uint8_t(m);
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
// neighbour references of the band cells (BandArgs::bandRef): position in the band, BAND_SKIP for a non-active neighbour,
// -2 - gridIndex for an active neighbour outside the band
__global__ void __launch_bounds__(BLOCK) k_band_ref(int32_t *bandRef, const int32_t *pos, const int32_t *bandIdx, const uint8_t *labels, int nBand,
						   int pitch, int64_t plane)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nBand) return;
    const int64_t i = bandIdx[k];
    const int64_t stride[6] = {-1, 1, -int64_t(pitch), int64_t(pitch), -plane, plane};
#pragma unroll
    for (int n = 0; n < 6; ++n)
    {
	const int64_t j = i + stride[n];
	const int l = labels[j];
	int ref = BAND_SKIP;
	if (l == L_INTERIOR || l == L_BOUNDARY)
	{
	    const int p = pos[j];
	    ref = p >= 0 ? p : int(-2 - j);
	}
	bandRef[int64_t(n) * nBand + k] = ref;
    }
}
// coefficient record of a BOUNDARY cell (Ops.h:208-255): c_n = 1 (INTERIOR nbr), w (BOUNDARY nbr), 0 otherwise;
// diag = sum of 1 (INTERIOR), w (BOUNDARY), w (DIRICHLET).  w == nullptr (coarse levels) means weight 1.
// wcode[k]: two bits per neighbour n (bits 2n, 2n+1): 0 = coefficient 0 (skip), 1 = coefficient exactly 1 (no multiply, no load),
// 2 = fractional (load it).  Most BOUNDARY cells have no fractional coefficient at all -- a ghost-fluid weight sits on a
// liquid/air face, whose neighbour is DIRICHLET and only enters the diagonal -- so the sweeps read 2 bytes instead of 48.
__device__ __forceinline__ unsigned coefCode(double cn) { return cn == 0.0 ? 0u : (cn == 1.0 ? 1u : 2u); }

__global__ void __launch_bounds__(BLOCK) k_band_coef(double *bcoef, unsigned short *wcode, const int32_t *bandIdx, int nBoundary, const uint8_t *labels,
						    const double *w0, const double *w1, const double *w2, int pitch, int64_t plane)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nBoundary) return;
    const int64_t i = bandIdx[k];
    const int64_t stride[3] = {1, pitch, plane};
    const double *w[3] = {w0, w1, w2};
    double diag = 0.0, wsum = 0.0;
    unsigned code = 0;
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
	wsum += wt;
	code |= coefCode(cn) << (2 * n);
	bcoef[int64_t(n) * nBoundary + k] = cn;
    }
    wcode[k] = (unsigned short)code;
    bcoef[int64_t(6) * nBoundary + k] = diag;
    // row 7: the plain sum of the six face weights -- the diagonal the reference's diagonal preconditioner inverts
    // (GFS.cpp:538-548); without weight grids there is no such sum and the operator's diagonal stands in
    bcoef[int64_t(7) * nBoundary + k] = w0 ? wsum : diag;
}

// The same records from SPARSE face weights: wf[n][k] = weight of face n (-x,+x,-y,+y,-z,+z) of BOUNDARY cell k, gathered on the
// host from the caller's weight grids (only BOUNDARY cells ever look at a face weight, so the full grids never travel).
__global__ void __launch_bounds__(BLOCK) k_band_coef_sparse(double *bcoef, unsigned short *wcode, const int32_t *bandIdx, int nBoundary, const uint8_t *labels,
								   const double *wf, int pitch, int64_t plane)
{
    const int k = blockIdx.x * BLOCK + threadIdx.x;
    if (k >= nBoundary) return;
    const int64_t i = bandIdx[k];
    const int64_t stride[3] = {1, pitch, plane};
    double diag = 0.0, wsum = 0.0;
    unsigned code = 0;
#pragma unroll
    for (int n = 0; n < 6; ++n)
    {
	const int axis = n >> 1;
	const int64_t j = (n & 1) ? i + stride[axis] : i - stride[axis];
	const int nl = labels[j];
	const double wt = wf[int64_t(n) * nBoundary + k];
	double cn = 0.0;
	if (nl == L_INTERIOR) { cn = 1.0; diag += 1.0; }
	else if (nl == L_BOUNDARY) { cn = wt; diag += wt; }
	else if (nl == L_DIRICHLET) { diag += wt; }
	wsum += wt;
	code |= coefCode(cn) << (2 * n);
	bcoef[int64_t(n) * nBoundary + k] = cn;
    }
    wcode[k] = (unsigned short)code;
    bcoef[int64_t(6) * nBoundary + k] = diag;
    bcoef[int64_t(7) * nBoundary + k] = wsum;  // GFS.cpp:538-548
}

// Diagonal preconditioner grid of GFS.cpp:493-560: 1/6 on INTERIOR cells, 1/(sum of the six face weights) on BOUNDARY cells,
// 0 elsewhere.  CTAs [0, nChunks) walk the active chunks, the rest the boundary records.
__global__ void __launch_bounds__(BLOCK) k_diag_inverse(double *d, const uint8_t *labels, const int32_t *chunks, int nChunks, int chunksPerPlane,
						       int64_t plane, int nz, const int32_t *bandIdx, const double *bcoef, int nBoundary)
{
    if (int(blockIdx.x) < nChunks)
    {
	const int c = chunks[blockIdx.x];
	const int zb = c / chunksPerPlane;
	const int64_t inPlane = int64_t(c - zb * chunksPerPlane) * CHUNK_CELLS + 2 * threadIdx.x;
	if (inPlane >= plane) return;
	for (int dz = 0; dz < CHUNK_Z; ++dz)
	{
	    const int z = zb * CHUNK_Z + dz;
	    if (z >= nz) break;
	    const int64_t i = int64_t(z) * plane + inPlane;
	    if (labels[i] == L_INTERIOR) d[i] = 1. / 6.;
	    if (labels[i + 1] == L_INTERIOR) d[i + 1] = 1. / 6.;
	}
    }
    else
    {
	const int k = (blockIdx.x - nChunks) * BLOCK + threadIdx.x;
	if (k < nBoundary) d[bandIdx[k]] = 1. / bcoef[int64_t(7) * nBoundary + k];
    }
}

// bounding rectangle of the ACTIVE cells of every z-plane: ext[z] = (x0, x1, y0, y1), half-open; x1 <= x0 for an empty plane.
// Host transfers of vector grids only move these rectangles (everything else in the box is 0 by the vector-grid invariant).
__global__ void __launch_bounds__(BLOCK) k_plane_extents(int4 *ext, const uint8_t *labels, int n0, int n1, int pitch, int64_t plane)
{
    __shared__ int sx0, sx1, sy0, sy1;
    if (threadIdx.x == 0) { sx0 = n0; sx1 = 0; sy0 = n1; sy1 = 0; }
    __syncthreads();
    const uint8_t *p = labels + int64_t(blockIdx.x) * plane;
    int x0 = n0, x1 = 0, y0 = n1, y1 = 0;
    for (int i = threadIdx.x; i < n1 * pitch; i += BLOCK)
    {
	const int l = p[i];
	if (l == L_INTERIOR || l == L_BOUNDARY)
	{
	    const int y = i / pitch, x = i - y * pitch;
	    x0 = min(x0, x); x1 = max(x1, x + 1); y0 = min(y0, y); y1 = max(y1, y + 1);
	}
    }
    x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
    y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
    if ((threadIdx.x & 31) == 0) { atomicMin(&sx0, x0); atomicMax(&sx1, x1); atomicMin(&sy0, y0); atomicMax(&sy1, y1); }
    __syncthreads();
    if (threadIdx.x == 0) ext[blockIdx.x] = make_int4(sx0, sx1, sy0, sy1);
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
