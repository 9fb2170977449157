// gmg_p2p.cuh -- z-slab communication over NVLink peer memory, written by our own kernels.
//
// NCCL costs ~25 us per operation inside the PCG graph (measured, profiles/), and an iteration needs eight of them, most
// moving a few megabytes or one double.  Here every rank owns an ARENA (cudaMalloc, exported with CUDA IPC and mapped by
// its peers at context-shard time) holding mailboxes and sequence flags.  An exchange is ONE kernel per rank:
//   push  : copy my boundary planes straight into the neighbours' mailboxes (remote stores over NVLink)
//   signal: __threadfence_system, then the last CTA stores the sequence number into the neighbours' flags
//   pull  : wait until my own flags carry the sequence number, copy my mailboxes into my halo planes
// Mailboxes are double-buffered by sequence parity: a rank can only start exchange n+1 after it received the neighbour's
// data of exchange n, which the neighbour sent after it had consumed exchange n-1 -- the slot that n+1 overwrites.
// The same scheme gathers the first replicated level's right-hand side and sums the CG scalars (fixed rank order, so
// every rank gets the bitwise same sum).
#pragma once

#include "gmg_common.cuh"

namespace gmg
{
constexpr int P2P_MAX_WORLD = 16;
constexpr int P2P_MAX_LEVELS = 4;
constexpr int P2P_CTAS = 512;              // every CTA must be resident at once (they wait on each other's flags): 4 per SM hold enough
                                           // remote stores in flight for NVLink (round 1's <= 69 CTAs kept ~280 KB in flight: ~110 GB/s)
// A rank that waits longer than this for a neighbour gives up loudly instead of hanging the GPU: the error word becomes
// sticky, the stale mailbox is NOT copied and the channel's sequence number is NOT advanced.  Wall-clock (%globaltimer), so
// legitimate skew between the ranks' host threads (first-use graph instantiation, a slow upload) is covered; GMG_P2P_TIMEOUT_S
// overrides the default at gmg_ctx_shard time.
__device__ unsigned long long g_p2pTimeoutNs = 120ull * 1000000000ull;
__device__ __forceinline__ unsigned long long globalTimerNs()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// layout of one rank's arena (identical on every rank)
struct P2pLayout
{
    size_t bytes = 0;
    size_t flags = 0;                      // unsigned long long [P2P_FLAG_COUNT]
    size_t scalars = 0;                    // double [2][P2P_MAX_WORLD]
    size_t halo[P2P_MAX_LEVELS][2][2];     // [level][from lower / from upper][slot] -> doubles
    size_t gather[2];                      // [slot] -> a whole grid of the first replicated level
};
// flag indices
constexpr int P2P_FLAG_HALO = 0;                                      // [level][dir][slot]
constexpr int P2P_FLAG_GATHER = P2P_MAX_LEVELS * 4;                   // [slot][rank]
constexpr int P2P_FLAG_SCALAR = P2P_FLAG_GATHER + 2 * P2P_MAX_WORLD;  // [slot][rank]
constexpr int P2P_FLAG_COUNT = P2P_FLAG_SCALAR + 2 * P2P_MAX_WORLD;

struct P2pState
{
    int rank = 0, world = 1;
    int generation = 0;                          // bumped whenever the arenas are re-allocated (older solvers' graphs point into freed memory)
    P2pLayout layout;                            // ONE layout per generation, shared by every solver of the context
    size_t haloCap[P2P_MAX_LEVELS] = {0, 0, 0, 0};  // bytes one slot of a level's halo boxes holds; capacities only grow
    size_t gatherCap = 0;
    char *arena = nullptr;                       // mine
    char *peer[P2P_MAX_WORLD] = {nullptr};       // peers' arenas mapped into this process (peer[rank] == arena)
    // local (not shared) counters: sequence numbers per channel, CTA tickets, error word
    unsigned long long *seq = nullptr;           // [P2P_MAX_LEVELS + 2]: halo per level, gather, scalar
    unsigned *tickets = nullptr;                 // [2]
    int *error = nullptr;
};

__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stReleaseSys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one thread polls; false after the timeout (or when an earlier exchange already failed: the error is sticky)
__device__ __forceinline__ bool pollFlag(const unsigned long long *flag, unsigned long long want, int *error)
{
    if (ldAcquireSys(flag) >= want) return true;
    if (*reinterpret_cast<volatile int *>(error)) return false;
    const unsigned long long t0 = globalTimerNs();
    unsigned spins = 0;
    while (ldAcquireSys(flag) < want)
    {
	if ((++spins & 1023u) == 0 && globalTimerNs() - t0 > g_p2pTimeoutNs) { atomicExch(error, 1); return false; }
    }
    return true;
}
// thread 0 polls, the CTA follows; returns whether the data arrived
__device__ __forceinline__ bool waitFlag(const unsigned long long *flag, unsigned long long want, int *error)
{
    __shared__ int arrived;
    if (threadIdx.x == 0) arrived = pollFlag(flag, want, error) ? 1 : 0;
    __syncthreads();
    const bool ok = arrived != 0;
    __syncthreads();
    return ok;
}
// all CTAs of the grid have passed this point once the returned value is true in the last one (one ticket per phase).
// The CTA barrier orders every thread's stores before thread 0's fence (the pattern of a cooperative grid sync), so one
// fence per CTA publishes the whole CTA's remote stores before the ticket -- and the flag that follows the last ticket.
template <bool SYSTEM>
__device__ __forceinline__ bool lastCta(unsigned *ticket)
{
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0)
    {
	if (SYSTEM) __threadfence_system();
	else __threadfence();
	last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    return last;
}
// `planes` planes, of each only the rows [rowLo, rowHi) -- the rows that hold an active cell anywhere in the level (the same
// on every rank: labels are replicated); everything else in a vector grid is 0 on both sides and stays so
__device__ __forceinline__ void copyRows(double *dst, const double *src, int planes, int64_t plane, int64_t segOff, int64_t segLen, bool srcIsMailbox)
{
    const int64_t n2 = segLen >> 1;  // the row pitch is even
    const int64_t total = int64_t(planes) * n2;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x)
    {
	const int64_t pl = i / n2, j = i - pl * n2;
	const double2 *s = reinterpret_cast<const double2 *>(src + pl * plane + segOff) + j;
	double2 *d = reinterpret_cast<double2 *>(dst + pl * plane + segOff) + j;
	*d = srcIsMailbox ? __ldcg(s) : *s;
    }
}
__device__ __forceinline__ void copyPlanes(double *dst, const double *src, int64_t n, bool srcIsMailbox)
{
    // n is a multiple of 2 (row pitch is a multiple of 16 doubles); mailboxes were written by another GPU: bypass L1
    const int64_t n2 = n >> 1;
    const double2 *s = reinterpret_cast<const double2 *>(src);
    double2 *d = reinterpret_cast<double2 *>(dst);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n2; i += int64_t(gridDim.x) * blockDim.x)
	d[i] = srcIsMailbox ? __ldcg(s + i) : s[i];
}

struct HaloP2pArgs
{
    double *grid;                 // plane 0 = first stored plane of the slab
    int64_t plane;
    int64_t segOff, segLen;       // doubles: offset and length of the active rows inside a plane (segLen even)
    int ownLo, ownHi, depth;
    int hasLower, hasUpper;
    double *toLower, *toUpper;    // [2 slots] REMOTE: the lower neighbour's "from upper" box, the upper neighbour's "from lower" box
    int64_t slotStride;           // doubles between the two slots of a box
    const double *fromLower, *fromUpper;  // LOCAL boxes
    unsigned long long *flagOnLower, *flagOnUpper;      // REMOTE [2 slots]
    const unsigned long long *myFlagLower, *myFlagUpper; // LOCAL [2 slots]
    unsigned long long *seq;      // LOCAL counter of this level's channel
    unsigned *tickets;
    int *error;
};

__global__ void __launch_bounds__(256) k_halo_p2p(const HaloP2pArgs a)
{
    const unsigned long long seq = *a.seq + 1;
    const int slot = int(seq & 1);
    if (a.hasLower) copyRows(a.toLower + slot * a.slotStride, a.grid + int64_t(a.ownLo) * a.plane, a.depth, a.plane, a.segOff, a.segLen, false);
    if (a.hasUpper) copyRows(a.toUpper + slot * a.slotStride, a.grid + int64_t(a.ownHi - a.depth) * a.plane, a.depth, a.plane, a.segOff, a.segLen, false);
    if (lastCta<true>(a.tickets))
    {
	if (threadIdx.x == 0)
	{
	    if (a.hasLower) stReleaseSys(a.flagOnLower + slot, seq);
	    if (a.hasUpper) stReleaseSys(a.flagOnUpper + slot, seq);
	}
    }
    if (a.hasLower)
    {
	if (waitFlag(a.myFlagLower + slot, seq, a.error))
	    copyRows(a.grid + int64_t(a.ownLo - a.depth) * a.plane, a.fromLower + slot * a.slotStride, a.depth, a.plane, a.segOff, a.segLen, true);
    }
    if (a.hasUpper)
    {
	if (waitFlag(a.myFlagUpper + slot, seq, a.error))
	    copyRows(a.grid + int64_t(a.ownHi) * a.plane, a.fromUpper + slot * a.slotStride, a.depth, a.plane, a.segOff, a.segLen, true);
    }
    if (lastCta<false>(a.tickets + 1))
    {
	if (threadIdx.x == 0)
	{
	    if (!*reinterpret_cast<volatile int *>(a.error)) *a.seq = seq;  // a failed exchange leaves the channel where it was
	    a.tickets[0] = 0;
	    a.tickets[1] = 0;
	}
    }
}

struct GatherP2pArgs
{
    double *grid;                       // the first replicated level's grid (whole box on every rank)
    int64_t plane;
    int rank, world;
    int lo[P2P_MAX_WORLD], hi[P2P_MAX_WORLD];  // planes each rank contributes
    double *peerBox[P2P_MAX_WORLD];     // REMOTE gather boxes [2 slots]
    int64_t slotStride;
    const double *myBox;                // LOCAL
    unsigned long long *flagOnPeer[P2P_MAX_WORLD];  // REMOTE [2 slots][world]
    const unsigned long long *myFlags;  // LOCAL [2 slots][world]
    unsigned long long *seq;
    unsigned *tickets;
    int *error;
};

__global__ void __launch_bounds__(256) k_gather_p2p(const GatherP2pArgs a)
{
    const unsigned long long seq = *a.seq + 1;
    const int slot = int(seq & 1);
    const int64_t off = int64_t(a.lo[a.rank]) * a.plane, n = int64_t(a.hi[a.rank] - a.lo[a.rank]) * a.plane;
    for (int r = 0; r < a.world; ++r)
	if (r != a.rank) copyPlanes(a.peerBox[r] + slot * a.slotStride + off, a.grid + off, n, false);
    if (lastCta<true>(a.tickets))
    {
	if (threadIdx.x < a.world && int(threadIdx.x) != a.rank) stReleaseSys(a.flagOnPeer[threadIdx.x] + slot * P2P_MAX_WORLD + a.rank, seq);
    }
    for (int r = 0; r < a.world; ++r)
    {
	if (r == a.rank) continue;
	if (!waitFlag(a.myFlags + slot * P2P_MAX_WORLD + r, seq, a.error)) continue;
	const int64_t o = int64_t(a.lo[r]) * a.plane;
	copyPlanes(a.grid + o, a.myBox + slot * a.slotStride + o, int64_t(a.hi[r] - a.lo[r]) * a.plane, true);
    }
    if (lastCta<false>(a.tickets + 1))
    {
	if (threadIdx.x == 0)
	{
	    if (!*reinterpret_cast<volatile int *>(a.error)) *a.seq = seq;  // a failed exchange leaves the channel where it was
	    a.tickets[0] = 0;
	    a.tickets[1] = 0;
	}
    }
}

struct ScalarP2pArgs
{
    double *value;                      // in: this rank's partial; out: the reduction over the ranks
    int rank, world, isMax;
    double *peerSlots[P2P_MAX_WORLD];   // REMOTE [2 slots][world]
    const double *mySlots;              // LOCAL
    unsigned long long *flagOnPeer[P2P_MAX_WORLD];
    const unsigned long long *myFlags;
    unsigned long long *seq;
    int *error;
};

// one CTA: thread r talks to rank r
__global__ void __launch_bounds__(32) k_scalar_p2p(const ScalarP2pArgs a)
{
    const unsigned long long seq = *a.seq + 1;
    const int slot = int(seq & 1);
    const int r = threadIdx.x;
    const double mine = *a.value;
    if (r < a.world && r != a.rank)
    {
	a.peerSlots[r][slot * P2P_MAX_WORLD + a.rank] = mine;
	__threadfence_system();
	stReleaseSys(a.flagOnPeer[r] + slot * P2P_MAX_WORLD + a.rank, seq);
	pollFlag(a.myFlags + slot * P2P_MAX_WORLD + r, seq, a.error);
    }
    __syncwarp();
    if (r == 0)
    {
	// fixed rank order: every rank computes the bitwise same result
	double acc = 0.0;
	for (int k = 0; k < a.world; ++k)
	{
	    const double v = (k == a.rank) ? mine : __ldcg(a.mySlots + slot * P2P_MAX_WORLD + k);
	    acc = a.isMax ? fmax(acc, v) : acc + v;
	}
	if (!*reinterpret_cast<volatile int *>(a.error))
	{
	    *a.value = acc;
	    *a.seq = seq;
	}
    }
}
} // namespace gmg
