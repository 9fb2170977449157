// gmg_common.cuh -- shared declarations of the B200 (sm_100a) MGPCG library.
//
// Data layout in HBM:
//   * every level stores only the cropped box of non-EXTERIOR cells plus a >=2-cell halo, as a dense
//     x-fastest array with a 128-byte aligned row pitch; expanded coordinates stay virtual.
//   * storage origins are even on every axis, so the two x-children of a coarse cell form an aligned
//     double2 and fine/coarse storage indices are related by cs = (fs >> 1) + shift.
//   * values are fp64 (the reference is fp64 end to end), labels are one byte.
//   * vector grids are exactly 0 on non-active cells; kernels only write active cells.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/gmg_b200.h"

namespace gmg
{
constexpr int L_INTERIOR = GMG_INTERIOR_CELL;
constexpr int L_EXTERIOR = GMG_EXTERIOR_CELL;
constexpr int L_DIRICHLET = GMG_DIRICHLET_CELL;
constexpr int L_BOUNDARY = GMG_BOUNDARY_CELL;

constexpr int BLOCK = 256;          // threads per CTA for every grid kernel
constexpr int CHUNK_CELLS = 512;    // cells of one z-plane a CTA covers (BLOCK threads x double2)
constexpr int CHUNK_Z = 4;          // z-planes a CTA walks
constexpr int MAX_COARSE = 4096;    // dense coarse-solve limit (unknowns)

// kernel classes for profiling / roofline accounting
enum KClass
{
    KC_JACOBI = 0,
    KC_APPLY,
    KC_RESIDUAL,
    KC_BAND,
    KC_RESTRICT,
    KC_PROLONG,
    KC_COARSE,
    KC_BLAS1,
    KC_REDUCE,
    KC_ZERO,
    KC_SETUP,
    KC_HALO,
    KC_GS,
    KC_COUNT
};

struct Geom
{
    int64_t res[3];  // expanded (virtual) resolution of this level
    int org[3];      // expanded coordinate of storage cell (0,0,0); even
    int n[3];        // storage extents
    int pitch;       // x pitch in elements (multiple of 16)
    int64_t plane;   // pitch * n[1]
    int64_t total;   // plane * n[2]
    int chunksPerPlane;
    int zBlocks;
};

// deep-halo depths of the z-slab sharding, in planes of the level they apply to
constexpr int HALO_X = 8;        // rhs / solution refresh depth per sharded level (7 smoothing sweeps + residual / prolongation)
constexpr int HALO_P = 9;        // CG search direction at level 0 (A p valid to depth 8 keeps r consistent without its own exchange)
constexpr int HALO_STORE0 = 10;  // stored halo planes at level 0 (even, >= HALO_P)
constexpr int HALO_STORE = 8;    // stored halo planes at the other sharded levels

struct Level
{
    Geom g;                      // LOCAL storage box: the rank's z-slab (+ stored halo) on a sharded level, else == gg
    Geom gg;                     // global storage box of the level
    bool sharded = false;
    int zOff = 0;                // local plane 0 is global storage plane zOff (even)
    int ownLo = 0, ownHi = 0;    // owned planes [ownLo, ownHi) in local storage coordinates
    int rowLo = 0, rowHi = 0;    // sharded levels: rows [rowLo, rowHi) of a plane hold every active cell of the level (all a halo exchange moves)
    uint8_t *labelsAlloc = nullptr; // the level's labels over the GLOBAL box (every rank holds them; 1 byte per cell)
    uint8_t *labels = nullptr;   // = labelsAlloc + zOff * plane
    uint8_t *flagsAlloc = nullptr;  // band flags over the GLOBAL box (gmg_kernels.cuh: SM_JACOBI_ZERO): bit0 in band, bit1 next to it
    uint8_t *bandFlags = nullptr;   // = flagsAlloc + zOff * plane
    uint8_t *nbrMaskAlloc = nullptr;  // neighbour masks over the GLOBAL box (k_band_nbr_mask): what the zero-aware interior sweep reads
    uint8_t *nbrMask = nullptr;       // = nbrMaskAlloc + zOff * plane
    int64_t nActive = 0, nInterior = 0; // over the local stored box
    int64_t nActiveGlobal = 0;
    // boundary band: [0,nBoundary) BOUNDARY cells, [nBoundary,nBand) INTERIOR cells of the band; linear order inside each part
    int nBoundary = 0, nBand = 0;
    char *bandSlab = nullptr;    // one allocation behind bandIdx / bandRef / bandV0 / bandV1 / bandB / bcoef (L2 persisting window)
    size_t bandSlabBytes = 0;
    int32_t *bandIdx = nullptr;  // [nBand] storage index
    int32_t *bandRef = nullptr;  // [6][nBand] neighbour reference (gmg_kernels.cuh: BandArgs::bandRef)
    bool hasWeights = false;     // level 0 built with face weights: BOUNDARY cells carry fractional coefficients
    double *bcoef = nullptr;     // [8][nBoundary]: coefficient on each of the 6 neighbours, the diagonal, the sum of the six face weights
    unsigned short *wcode = nullptr;  // [nBoundary] two bits per neighbour: coefficient 0 / exactly 1 / fractional (gmg_kernels.cuh: coefCode)
    double *bandV0 = nullptr, *bandV1 = nullptr, *bandB = nullptr;
    // CTAs of the full-grid kernels: chunks holding at least one INTERIOR cell / one active cell
    int nChunksInterior = 0, nChunksActive = 0;
    int32_t *chunksInterior = nullptr, *chunksActive = nullptr;
    // TMA path (k_stencil_tma): 64 x 8 x 4 bricks holding an INTERIOR cell (linear brick ids, x fastest); null = plain-load kernels
    int32_t *bricks = nullptr;
    int nBricks = 0, bricksX = 0, bricksY = 0;
    int32_t *bricksActive = nullptr;  // the same bricks holding an ACTIVE cell (TMA prolongation, k_prolong_tma)
    int nBricksActive = 0;
    int32_t *cbricks = nullptr;  // 32 x 4 x 2 bricks of THIS level's cells holding an active cell: TMA restriction into this level (k_restrict_tma)
    int nCBricks = 0, cbricksX = 0, cbricksY = 0;
    // V-cycle grids (level 0 uses caller grids for x and b)
    double *x = nullptr, *xAlt = nullptr, *b = nullptr, *r = nullptr;
    // mixed precision (gmg_solver_options::mixed_precision): the V-cycle's grids and compact band arrays once more in fp32
    float *x32 = nullptr, *xAlt32 = nullptr, *b32 = nullptr, *r32 = nullptr;
    float *bandV0f = nullptr, *bandV1f = nullptr, *bandBf = nullptr;
    void *smoothArgs = nullptr;  // gmg::ClusterSmoothArgs: this level smooths inside one thread-block cluster (gmg_cluster.cuh: k_cluster_smooth)
    void *smoothSlab = nullptr;  // its tables
    size_t smoothSmem = 0;
    void *tileArgs = nullptr;    // gmg::BandTileArgs: the band sweep groups of this level run as tiles with ring halos (gmg_band_tiles.cuh: k_band_tile)
    void *tileSlab = nullptr;    // its tables
    size_t tileSmem = 0;
    int nTiles = 0;
    bool tilesCoResident = false;  // every tile fits on the device at once: the groups that start from a non-zero grid may use them too
    int shift[3] = {0, 0, 0};    // coarse storage = (this level's storage >> 1) + shift   (to level+1)
    // tiled Gauss-Seidel (only built when the solver uses it): 16^3 tiles of the EXPANDED grid laid over the storage box
    int32_t *gsTiles[2] = {nullptr, nullptr};  // [0] even, [1] odd tiles holding an active cell (linear tile ids)
    int nGsTiles[2] = {0, 0};
    int gsTilesX = 0, gsTilesY = 0, gsOff[3] = {0, 0, 0};
    int32_t *bpos = nullptr;     // grid: boundary-record index of BOUNDARY cells
};

// host transfers of level-0 vector grids: groups of consecutive z-planes with the bounding rectangle of their active cells
// (local storage coordinates, half-open)
struct IoGroup
{
    int z0, z1, x0, x1, y0, y1;
};

struct TmaMap { alignas(64) unsigned char bytes[128]; };  // an opaque CUtensorMap (encoded in gmg_b200.cu, consumed by k_stencil_tma)

struct ProfileRec
{
    int klass;
    int level;
    double bytes;
    cudaEvent_t e0, e1;
    int n;  // launches inside the bracket (1, or the sweeps of a band sweep group in profiling mode 2)
};

} // namespace gmg

struct gmg_solver;
struct gmg_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int64_t launches = 0;
    bool profiling = false;
    bool profileGroups = false;   // gmg_profile_enable(ctx, 2): ONE event pair around the back-to-back sweeps of a band sweep group instead of one per
				  // sweep -- the event-record nodes cost ~5 us per bracket and break the prologue overlap between the launches they separate
    bool scopeMuted = false;      // inside such a group bracket: the launches' own brackets only count
    bool capturing = false;       // inside a stream capture: profiling events become external event-record nodes
    std::vector<gmg::ProfileRec> recs;
    std::vector<cudaEvent_t> eventPool;
    // [0] = all levels, [1] = launches on level 0 only (the fine level, where the roofline is quoted)
    double classMs[2][gmg::KC_COUNT] = {{0}};
    int64_t classLaunches[2][gmg::KC_COUNT] = {{0}};
    double classBytes[2][gmg::KC_COUNT] = {{0}};
    double levelMs[16][gmg::KC_COUNT] = {{0}};      // the same device times split by multigrid level
    int64_t levelLaunches[16][gmg::KC_COUNT] = {{0}};
    int curLevel = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int smCount = 148;
    size_t persistBytes = 0;      // L2 set aside for persisting lines (0 = off)
    size_t maxWindowBytes = 0;
    gmg_solver *windowOwner = nullptr;  // solver whose level-0 band slab the stream's access-policy window covers
    // sharding (z-slabs); world == 1 means single GPU
    int rank = 0, world = 1;
    void *nccl = nullptr;         // ncclComm_t
    int64_t commOps = 0;          // communication operations enqueued since the last launch-count reset
    void *p2p = nullptr;          // gmg::P2pState: peer-memory mailboxes (gmg_p2p.cuh); null = NCCL for every exchange
    bool p2pDisabled = false;
    int stencilSlots = 0;          // resident CTAs of k_stencil_loop
    int bandSlots[2] = {0, 0};     // resident CTAs of the band sweep kernel with two / three cells per thread
    int smoothClusterSize = 0;     // CTAs of the one-level smoothing cluster (k_cluster_smooth): 0 = not probed, -1 = refused
    int clusterSize = 0;           // CTAs of the coarse-cycle cluster: 0 = not probed yet, -1 = cluster launch refused
    bool deviceLoopBroken = false; // the driver refused the conditional-graph PCG loop once: host loop from then on
    int p2pGenerations = 0;
    std::vector<gmg_solver *> solvers;  // live solvers of this context (their cached graphs are dropped when the arenas are re-mapped)
    // reduction scratch
    double *partials = nullptr;   // [maxPartials]
    std::vector<void *> retired;  // outgrown scratch buffers that cached graphs of live solvers may still reference
    unsigned *ticket = nullptr;
    void *groupBarrier = nullptr;  // gmg::GroupBarrier of the persistent band sweep groups (k_band_group)
    int bandGroupCtas[2] = {0, 0}; // co-resident CTAs of k_band_group<unweighted / weighted> on this device
    double *scalars = nullptr;    // device scalars (see Scalars)
    double *hostScalars = nullptr; // pinned mirror
    int maxPartials = 0;
    // pinned double buffer for host<->device box transfers
    void *pin[2] = {nullptr, nullptr};
    size_t pinCap = 0;
    cudaEvent_t pinEv[2] = {nullptr, nullptr};
};

struct gmg_solver
{
    gmg_ctx *ctx = nullptr;
    gmg_solver_options opt;
    int levels = 0;
    std::vector<gmg::Level> lv;
    int64_t hostBounds[6] = {0, 0, 0, 0, 0, 0};  // expanded [lo, hi) of the non-EXTERIOR cells at level 0: host arrays are only read inside
    // coarsest direct solve
    int nCoarse = 0;
    int32_t *coarseIdx = nullptr; // [nCoarse] storage index at the coarsest level
    double *coarseInv = nullptr;  // [nCoarse][nCoarse] row-major inverse
    // z-slab sharding: levels [0, shardLevels) are slabs, the rest replicated on every rank
    int shardLevels = 0;
    int p2pGeneration = -1;       // generation of the context's peer-memory arenas this solver was built against
    std::vector<int> gatherLo, gatherHi; // per rank: planes of the first replicated level it restricts into
    // compact coarse sub-V-cycle: levels [fusedFirst, levels-1] in one shared-memory CTA (-1 = off)
    int fusedFirst = -1;
    void *compactArgs = nullptr;  // host copy of the kernel's CompactArgs
    void *compactBlob = nullptr;  // device tables behind it
    size_t compactSmem = 0;
    // the same levels in one thread-block cluster with distributed shared memory (gmg_cluster.cuh); preferred when available
    void *clusterArgs = nullptr;  // host copy of the kernel's ClusterArgs
    void *clusterSlab = nullptr;  // device tables behind it
    size_t clusterSmem = 0;
    // PCG work grids (level 0)
    double *pcgR = nullptr, *pcgP = nullptr, *pcgZ = nullptr, *pcgT = nullptr, *pcgX = nullptr, *pcgB = nullptr;
    double *diagInv = nullptr;           // diagonal preconditioner grid (GFS.cpp:487-560), built on first use
    void *pcgLoop = nullptr;             // device PcgLoopState + residual history of the device-side PCG loop
    void *pcgLoopHost = nullptr;         // pinned mirror
    double setupMs = 0;
    int64_t pcgSolves = 0;               // solves run on this solver so far
    std::vector<gmg::IoGroup> ioGroups;  // empty = move the whole box
    int64_t ioCells = 0;                 // cells one grid transfer moves
    // CUDA-graph cache: a V-cycle is ~100 dependent launches, most of them on tiny coarse levels, so the
    // launch sequence is captured once per (kind, x, b, flag) and replayed
    struct GraphEntry
    {
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t exec = nullptr;
	int64_t kernels = 0, comms = 0;
	std::vector<gmg::ProfileRec> recs; // profiled variant: per-launch event pairs living inside the graph
    };
    std::map<std::tuple<int, const void *, const void *, int>, GraphEntry> graphs;
    std::map<std::pair<int, const double *>, gmg::TmaMap> tensorMaps;  // (level, grid) -> tensor map of the TMA stencil kernels
    int tmaMask = 23;             // which full-grid kernels take the TMA-staged variant on big levels (GMG_TMA, see tmaMode)
    bool useGraphs = true;
    bool bandGroups = false;      // a group of band sweeps as ONE co-resident launch with grid barriers: measured SLOWER than a launch per sweep
				  // (profiles/r02_ab_switches.md: 256^3 solve 11.5 vs 10.3 ms; a barrier over ~900 CTAs costs more than a kernel
				  // boundary with its prologue overlapped), so it is opt-in (GMG_BAND_GROUPS=1)
    int stencilLoop = 0;          // persistent full-grid stencil kernels (k_stencil_loop): bit 0 Jacobi, bit 1 residual, bit 2 apply, bit 3 zero-aware Jacobi (GMG_STENCIL_LOOP)
    int stencilBatch = 0;         // full-grid stencil kernels with the loads of 2 / 4 planes issued together (k_stencil_b); 0 = one plane at a time (GMG_STENCIL_BATCH)
    int stencilCap = -1;          // full-grid stencil kernels capped at 40 registers: bit 0 Jacobi, bit 1 residual, bit 2 apply; -1 = by level size (GMG_STENCIL_CAP)
    int bandPerThread = 0;        // cells per thread of the band sweep kernels: 0 = picked per level (launchBand), 2 / 3 = forced (GMG_BAND_PER_THREAD)
    bool bandTiles = false;       // band sweep groups as ring-halo tiles, one launch per group (k_band_tile): measured SLOWER (GMG_BAND_TILES=1 enables)
    bool bandResident = false;    // a group of band sweeps as one launch with every cell's metadata on chip (k_band_resident): measured SLOWER too
				  // (11.04 vs 10.22 ms: with the prologues overlapped a kernel boundary costs what a 296-CTA barrier costs, ~2.5-3 us,
				  // and two 512-thread CTAs per SM gather with less parallelism than the sweep kernels); opt-in, GMG_BAND_RESIDENT=1
    int64_t bandGroupMax = 0;     // GMG_BAND_GROUP_MAX: the two one-launch group variants above only on levels whose band has at most this many cells
				  // (0 = no limit): a barrier over a few dozen CTAs is cheaper than one over 296 or ~900
    bool zeroAware = true;        // zero-aware down-stroke (no zero fill, SM_JACOBI_ZERO); GMG_ZERO_AWARE=0 at creation restores the fill
};

struct gmg_grid
{
    gmg_solver *solver = nullptr;
    int level = 0;
    double *d = nullptr;
    int64_t plane = 0; // guard-plane size: the allocation starts at d - plane
};

namespace gmg
{
void setError(const std::string &msg);
int cudaFail(cudaError_t e, const char *what, const char *file, int line);

#define GMG_CUDA(call)                                                           \
    do                                                                           \
    {                                                                            \
	cudaError_t _e = (call);                                                 \
	if (_e != cudaSuccess) return gmg::cudaFail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define GMG_TRY(call)                    \
    do                                   \
    {                                    \
	int _s = (call);                 \
	if (_s != GMG_OK) return _s;     \
    } while (0)

// profiling bracket around one kernel launch
struct LaunchScope
{
    gmg_ctx *ctx;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int klass;
    double bytes;
    int n;       // launches this bracket stands for; 0 = a group bracket that was not opened (profiling mode != 2)
    bool group;  // a group bracket: does not count as a launch itself, mutes the brackets of the launches inside it
    bool live = false;
    LaunchScope(gmg_ctx *c, int k, double b, int launches = 1, bool isGroup = false);
    ~LaunchScope();
};
#define GMG_LAUNCH(ctx, klass, bytes) gmg::LaunchScope _scope_##__LINE__(ctx, klass, bytes)

} // namespace gmg
