// gmg_b200.cu -- host side of the B200-native MGPCG library: C ABI (include/gmg_b200.h), solver
// construction (labels, bands, coarse factor), V-cycle and PCG drivers.  No CPU fallback: every
// entry point needs a CUDA device and fails with GMG_ERR_CUDA otherwise.
#include <cuda.h>  // CUtensorMap and its enums only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <thread>

#include <dlfcn.h>

#include "gmg_kernels.cuh"
#include "gmg_cluster.cuh"
#include "gmg_band_tiles.cuh"
#include "gmg_frontend.cuh"
#include "gmg_nccl.h"
#include "gmg_p2p.cuh"

using namespace gmg;

// ====================================================================================================
// errors, launch accounting
// ====================================================================================================
namespace
{
thread_local std::string g_lastError;
const char *kClassNames[KC_COUNT] = {"jacobi_interior", "apply_poisson", "residual", "band_jacobi", "restrict", "prolong_add",
				     "coarse_solve",    "blas1",         "reduce",   "zero_fill",   "setup",    "halo_exchange",
				     "gauss_seidel"};
inline int64_t divUp(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int floorDiv2(int64_t v) { return int(v >= 0 ? v / 2 : -((-v + 1) / 2)); }
inline int ceilDiv2(int64_t v) { return int(v >= 0 ? (v + 1) / 2 : -((-v) / 2)); }
double nowMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
} // namespace

namespace gmg
{
void setError(const std::string &msg) { g_lastError = msg; }
int cudaFail(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
    g_lastError = buf;
    return GMG_ERR_CUDA;
}
LaunchScope::LaunchScope(gmg_ctx *c, int k, double b, int launches, bool isGroup) : ctx(c), klass(k), bytes(b), n(launches), group(isGroup)
{
    if (!group)
    {
	if (k == KC_HALO) ++ctx->commOps;  // NCCL operations are not this library's kernels
	else ++ctx->launches;
    }
    if (!ctx->profiling || ctx->scopeMuted) return;
    if (group)
    {
	if (!ctx->profileGroups || n < 2) return;
	ctx->scopeMuted = true;
    }
    live = true;
    auto get = [&]() {
	cudaEvent_t e;
	if (!ctx->eventPool.empty()) { e = ctx->eventPool.back(); ctx->eventPool.pop_back(); }
	else cudaEventCreate(&e);
	return e;
    };
    e0 = get();
    e1 = get();
    if (ctx->capturing) cudaEventRecordWithFlags(e0, ctx->stream, cudaEventRecordExternal);
    else cudaEventRecord(e0, ctx->stream);
}
LaunchScope::~LaunchScope()
{
    if (!live) return;
    if (group) ctx->scopeMuted = false;
    if (ctx->capturing) cudaEventRecordWithFlags(e1, ctx->stream, cudaEventRecordExternal);
    else cudaEventRecord(e1, ctx->stream);
    ctx->recs.push_back({klass, ctx->curLevel, bytes, e0, e1, n});
}
} // namespace gmg


// Stream-ordered allocation from the device's default memory pool with an unlimited release threshold: a solver is
// built and torn down every simulation frame, so freed blocks stay cached in the pool instead of going back to the driver.
thread_local cudaStream_t g_stream = nullptr;
static cudaError_t enterCtx(gmg_ctx *ctx)
{
    g_stream = ctx->stream;
    return cudaSetDevice(ctx->device);
}
static bool usePool()
{
    static const bool v = [] { const char *e = getenv("GMG_POOL"); return !(e && e[0] == '0'); }();
    return v;
}
template <typename T>
static cudaError_t devMalloc(T **p, size_t bytes)
{
    if (!usePool()) return cudaMalloc(reinterpret_cast<void **>(p), bytes ? bytes : 16);
    return cudaMallocAsync(reinterpret_cast<void **>(p), bytes ? bytes : 16, g_stream);
}
static cudaError_t devFree(void *p)
{
    if (!p) return cudaSuccess;
    if (!usePool()) return cudaFree(p);
    return cudaFreeAsync(p, g_stream);
}

// Vector grids carry one zero guard plane below and above the stored box: on a z-slab the first / last stored plane has
// no EXTERIOR halo of its own, and the stencils of its cells reach one plane out.
static int allocGrid(double **p, const Geom &g)
{
    double *base = nullptr;
    const int64_t n = g.total + 2 * g.plane;
    GMG_CUDA(devMalloc(&base, sizeof(double) * n));
    GMG_CUDA(cudaMemsetAsync(base, 0, sizeof(double) * n, g_stream));
    *p = base + g.plane;
    return GMG_OK;
}
static void freeGrid(double *p, const Geom &g)
{
    if (p) devFree(p - g.plane);
}
static int allocGrid32(float **p, const Geom &g)
{
    float *base = nullptr;
    const int64_t n = g.total + 2 * g.plane;
    GMG_CUDA(devMalloc(&base, sizeof(float) * n));
    GMG_CUDA(cudaMemsetAsync(base, 0, sizeof(float) * n, g_stream));
    *p = base + g.plane;
    return GMG_OK;
}
static void freeGrid32(float *p, const Geom &g)
{
    if (p) devFree(p - g.plane);
}

// Programmatic dependent launch: a V-cycle is a chain of ~60 dependent kernels, most of them a few microseconds long, so the
// launch-to-launch gap matters as much as the kernels.  Every hot kernel starts with griddepcontrol.launch_dependents +
// griddepcontrol.wait (gmg_kernels.cuh: pdlEnter), and is launched with the programmatic-stream-serialization attribute:
// the next kernel's CTAs are set up while the current one still runs and only wait for its memory to be flushed.
static bool usePdl()
{
    static const bool v = [] { const char *e = getenv("GMG_PDL"); return !(e && e[0] == '0'); }();
    return v;
}
template <typename... KArgs, typename... Args>
static cudaError_t launchK(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = usePdl() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#define TRACE(msg) do { if (getenv("GMG_TRACE")) { fprintf(stderr, "[gmg] %s:%d %s\n", __func__, __LINE__, msg); fflush(stderr); } } while (0)

static void p2pRelease(gmg_ctx *ctx);
static int p2pCheckError(gmg_ctx *ctx);
static double *scalarPtr(gmg_solver *s, size_t offset);
static void dropGraphs(gmg_solver *s)
{
    for (auto &g : s->graphs) { cudaGraphExecDestroy(g.second.exec); cudaGraphDestroy(g.second.graph); }
    s->graphs.clear();
}

static int invalid(const char *msg)
{
    setError(msg);
    return GMG_ERR_INVALID;
}

static void accumulateRecs(gmg_ctx *ctx, const std::vector<ProfileRec> &recs, bool recycle)
{
    for (auto &r : recs)
    {
	float ms = 0;
	cudaEventElapsedTime(&ms, r.e0, r.e1);
	for (int f = 0; f < (r.level == 0 ? 2 : 1); ++f)
	{
	    ctx->classMs[f][r.klass] += ms;
	    ctx->classLaunches[f][r.klass] += r.n;
	    ctx->classBytes[f][r.klass] += r.bytes;
	}
	if (r.level >= 0 && r.level < 16) { ctx->levelMs[r.level][r.klass] += ms; ctx->levelLaunches[r.level][r.klass] += r.n; }
	if (recycle)
	{
	    ctx->eventPool.push_back(r.e0);
	    ctx->eventPool.push_back(r.e1);
	}
    }
}
static void flushProfile(gmg_ctx *ctx)
{
    if (ctx->recs.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    accumulateRecs(ctx, ctx->recs, true);
    ctx->recs.clear();
}

static int ensureScratch(gmg_ctx *ctx, int nPartials)
{
    if (nPartials <= ctx->maxPartials) return GMG_OK;
    // Reduction kernels captured in the cached graphs of solvers that are still alive hold the old pointer as a kernel
    // argument: the outgrown buffer is retired (freed with the context), never freed here.  Solvers of one context run
    // one after the other on its stream, so sharing a buffer between them is safe; an old graph keeps using the smaller
    // buffer, which was sized for that solver's own grids.
    if (ctx->partials) ctx->retired.push_back(ctx->partials);
    ctx->partials = nullptr;
    ctx->maxPartials = nPartials + 1024;
    GMG_CUDA(cudaMalloc(&ctx->partials, sizeof(double) * ctx->maxPartials));
    return GMG_OK;
}

// ====================================================================================================
// context
// ====================================================================================================
extern "C" const char *gmg_last_error(void) { return g_lastError.c_str(); }
extern "C" int gmg_version(void) { return 100; }

extern "C" int gmg_ctx_destroy(gmg_ctx *ctx);
extern "C" int gmg_ctx_create(int device, void *stream, gmg_ctx **out)
{
    if (!out) return invalid("gmg_ctx_create: out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
    {
	setError("gmg_ctx_create: no CUDA device available (this library has no CPU fallback)");
	return GMG_ERR_CUDA;
    }
    if (device < 0 || device >= count) return invalid("gmg_ctx_create: device ordinal out of range");
    GMG_CUDA(cudaSetDevice(device));
    gmg_ctx *ctx = new gmg_ctx;
    ctx->device = device;
    // every failure below goes through gmg_ctx_destroy, which frees whatever was created so far
    auto build = [&]() -> int {
	if (stream) ctx->stream = static_cast<cudaStream_t>(stream);
	else
	{
	    GMG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	    ctx->ownStream = true;
	}
	g_stream = ctx->stream;
	{
	    cudaMemPool_t pool;
	    GMG_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
	    uint64_t threshold = UINT64_MAX;
	    GMG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
	}
	cudaDeviceProp prop;
	GMG_CUDA(cudaGetDeviceProperties(&prop, device));
	ctx->smCount = prop.multiProcessorCount;
	{
	    // up to half of L2 set aside for persisting lines (the band slab of the solver in use, applyBandWindow)
	    const char *e = getenv("GMG_L2_PERSIST");
	    if (e && e[0] != '0' && prop.persistingL2CacheMaxSize > 0)
	    {
		size_t want = std::min<size_t>(size_t(prop.persistingL2CacheMaxSize), size_t(prop.l2CacheSize) / 2);
		if (e && atoi(e) > 1) want = std::min<size_t>(size_t(prop.persistingL2CacheMaxSize), size_t(atoi(e)) << 20);
		if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
		{
		    ctx->persistBytes = want;
		    ctx->maxWindowBytes = size_t(prop.accessPolicyMaxWindowSize);
		}
		else cudaGetLastError();
	    }
	}
	GMG_CUDA(cudaMalloc(&ctx->groupBarrier, sizeof(GroupBarrier)));
	GMG_CUDA(cudaMemset(ctx->groupBarrier, 0, sizeof(GroupBarrier)));
	GMG_CUDA(cudaMalloc(&ctx->ticket, sizeof(unsigned)));
	GMG_CUDA(cudaMemset(ctx->ticket, 0, sizeof(unsigned)));
	GMG_CUDA(cudaMalloc(&ctx->scalars, sizeof(Scalars)));
	GMG_CUDA(cudaMemset(ctx->scalars, 0, sizeof(Scalars)));
	GMG_CUDA(cudaMallocHost(&ctx->hostScalars, sizeof(Scalars)));
	GMG_CUDA(cudaEventCreate(&ctx->t0));
	GMG_CUDA(cudaEventCreate(&ctx->t1));
	GMG_TRY(ensureScratch(ctx, 4096));
	if (const char *e = getenv("GMG_PDL_PREFETCH"))
	{
	    const int waitFirst = e[0] == '0' ? 1 : 0;
	    GMG_CUDA(cudaMemcpyToSymbol(c_pdlWaitFirst, &waitFirst, sizeof(int)));
	}
	return GMG_OK;
    };
    const int st = build();
    if (st != GMG_OK)
    {
	const std::string why = g_lastError;
	gmg_ctx_destroy(ctx);
	g_lastError = why;
	return st;
    }
    *out = ctx;
    return GMG_OK;
}

extern "C" int gmg_ctx_destroy(gmg_ctx *ctx)
{
    if (!ctx) return GMG_OK;
    enterCtx(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    flushProfile(ctx);
    for (auto e : ctx->eventPool) cudaEventDestroy(e);
    p2pRelease(ctx);
    if (ctx->nccl)
	if (const NcclApi *api = ncclApi(nullptr)) api->CommDestroy(static_cast<NcclComm>(ctx->nccl));
    cudaFree(ctx->partials);
    for (void *p : ctx->retired) cudaFree(p);
    cudaFree(ctx->ticket);
    cudaFree(ctx->groupBarrier);
    cudaFree(ctx->scalars);
    cudaFreeHost(ctx->hostScalars);
    for (int i = 0; i < 2; ++i)
	if (ctx->pin[i]) { cudaFreeHost(ctx->pin[i]); cudaEventDestroy(ctx->pinEv[i]); }
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
    return GMG_OK;
}

extern "C" int gmg_ctx_synchronize(gmg_ctx *ctx)
{
    if (!ctx) return invalid("null ctx");
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return p2pCheckError(ctx);
}

// ---- NCCL, bound at run time (gmg_nccl.h) ----------------------------------------------------------
namespace gmg
{
const NcclApi *ncclApi(const char **why)
{
    static NcclApi api;
    static std::string err;
    static std::once_flag once;
    std::call_once(once, [] {
	const char *names[] = {getenv("GMG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	for (const char *n : names)
	{
	    if (!n || !n[0]) continue;
	    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
	    if (api.lib) break;
	    err = dlerror();
	}
	if (!api.lib) return;
	bool ok = true;
	auto sym = [&](const char *name) { void *p = dlsym(api.lib, name); if (!p) { ok = false; err = std::string("missing symbol ") + name; } return p; };
	api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
	api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
	api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
	api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
	api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
	api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
	api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
	api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
	api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
	api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
	api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
	if (!ok) { dlclose(api.lib); api.lib = nullptr; }
    });
    if (!api.lib) { if (why) *why = err.c_str(); return nullptr; }
    return &api;
}
} // namespace gmg

static int ncclFail(int r, const char *what)
{
    const NcclApi *api = ncclApi(nullptr);
    char buf[512];
    snprintf(buf, sizeof(buf), "NCCL error %d (%s) in %s", r, api ? api->GetErrorString(r) : "?", what);
    setError(buf);
    return GMG_ERR_COMM;
}
#define GMG_NCCL(call)                                        \
    do                                                        \
    {                                                         \
	int _r = (call);                                      \
	if (_r != NCCL_SUCCESS) return ncclFail(_r, #call);   \
    } while (0)

extern "C" int gmg_nccl_unique_id(void *out128)
{
    if (!out128) return invalid("gmg_nccl_unique_id: null argument");
    const char *why = nullptr;
    const NcclApi *api = ncclApi(&why);
    if (!api) { setError(std::string("gmg_nccl_unique_id: cannot load libnccl.so.2: ") + (why ? why : "")); return GMG_ERR_COMM; }
    GMG_NCCL(api->GetUniqueId(static_cast<NcclUniqueId *>(out128)));
    return GMG_OK;
}

extern "C" int gmg_ctx_shard(gmg_ctx *ctx, int rank, int world, const void *ncclUniqueId)
{
    if (!ctx) return invalid("gmg_ctx_shard: null ctx");
    if (world < 1 || rank < 0 || rank >= world) return invalid("gmg_ctx_shard: rank/world out of range");
    if (ctx->nccl) return invalid("gmg_ctx_shard: context is already sharded");
    if (world == 1) return GMG_OK;
    if (!ncclUniqueId) return invalid("gmg_ctx_shard: null unique id");
    const char *why = nullptr;
    const NcclApi *api = ncclApi(&why);
    if (!api) { setError(std::string("gmg_ctx_shard: cannot load libnccl.so.2: ") + (why ? why : "")); return GMG_ERR_COMM; }
    GMG_CUDA(enterCtx(ctx));
    NcclUniqueId id;
    std::memcpy(&id, ncclUniqueId, sizeof(id));
    NcclComm comm = nullptr;
    GMG_NCCL(api->CommInitRank(&comm, world, id, rank));
    ctx->nccl = comm;
    ctx->rank = rank;
    ctx->world = world;
    // one tiny all-reduce now: connection setup happens outside any later stream capture
    GMG_NCCL(api->AllReduce(ctx->scalars, ctx->scalars, 1, NCCL_FLOAT64, NCCL_SUM, comm, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(cudaMemsetAsync(ctx->scalars, 0, sizeof(Scalars), ctx->stream));
    if (const char *e = getenv("GMG_P2P_TIMEOUT_S"))
    {
	const unsigned long long ns = (unsigned long long)(std::max(1.0, atof(e)) * 1e9);
	GMG_CUDA(cudaMemcpyToSymbol(g_p2pTimeoutNs, &ns, sizeof(ns)));
    }
    return GMG_OK;
}
extern "C" int gmg_ctx_rank(gmg_ctx *ctx, int *rank, int *world)
{
    if (!ctx) return invalid("null ctx");
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    return GMG_OK;
}

extern "C" int gmg_launch_count(gmg_ctx *ctx, int64_t *count, int reset)
{
    if (!ctx) return invalid("null ctx");
    if (count) *count = ctx->launches;
    if (reset) { ctx->launches = 0; ctx->commOps = 0; }
    return GMG_OK;
}
extern "C" int gmg_comm_count(gmg_ctx *ctx, int64_t *count)
{
    if (!ctx || !count) return invalid("null argument");
    *count = ctx->commOps;
    return GMG_OK;
}
extern "C" int gmg_timer_begin(gmg_ctx *ctx)
{
    GMG_CUDA(cudaEventRecord(ctx->t0, ctx->stream));
    return GMG_OK;
}
extern "C" int gmg_timer_end(gmg_ctx *ctx, double *ms)
{
    GMG_CUDA(cudaEventRecord(ctx->t1, ctx->stream));
    GMG_CUDA(cudaEventSynchronize(ctx->t1));
    float f = 0;
    GMG_CUDA(cudaEventElapsedTime(&f, ctx->t0, ctx->t1));
    if (ms) *ms = f;
    return GMG_OK;
}
extern "C" int gmg_profile_enable(gmg_ctx *ctx, int on)
{
    flushProfile(ctx);
    ctx->profiling = on != 0;
    ctx->profileGroups = on == 2;
    return GMG_OK;
}
extern "C" int gmg_kernel_class_count(void) { return KC_COUNT; }
extern "C" const char *gmg_kernel_class_name(int i) { return (i >= 0 && i < KC_COUNT) ? kClassNames[i] : ""; }
extern "C" int gmg_profile_get(gmg_ctx *ctx, int klass, int fineLevelOnly, double *ms, int64_t *launches, double *bytes)
{
    if (klass < 0 || klass >= KC_COUNT) return invalid("bad kernel class");
    flushProfile(ctx);
    const int f = fineLevelOnly ? 1 : 0;
    if (ms) *ms = ctx->classMs[f][klass];
    if (launches) *launches = ctx->classLaunches[f][klass];
    if (bytes) *bytes = ctx->classBytes[f][klass];
    return GMG_OK;
}
extern "C" int gmg_profile_reset(gmg_ctx *ctx)
{
    flushProfile(ctx);
    for (int f = 0; f < 2; ++f)
	for (int i = 0; i < KC_COUNT; ++i) { ctx->classMs[f][i] = 0; ctx->classLaunches[f][i] = 0; ctx->classBytes[f][i] = 0; }
    for (int l = 0; l < 16; ++l)
	for (int i = 0; i < KC_COUNT; ++i) { ctx->levelMs[l][i] = 0; ctx->levelLaunches[l][i] = 0; }
    return GMG_OK;
}
extern "C" int gmg_profile_get_level(gmg_ctx *ctx, int klass, int level, double *ms, int64_t *launches)
{
    if (klass < 0 || klass >= KC_COUNT || level < 0 || level >= 16) return invalid("bad kernel class / level");
    flushProfile(ctx);
    if (ms) *ms = ctx->levelMs[level][klass];
    if (launches) *launches = ctx->levelLaunches[level][klass];
    return GMG_OK;
}

// ====================================================================================================
// geometry helpers
// ====================================================================================================
static BoxArgs boxArgs(const Geom &g)
{
    BoxArgs b;
    for (int a = 0; a < 3; ++a) { b.n[a] = g.n[a]; b.org[a] = g.org[a]; b.res[a] = g.res[a]; }
    b.pitch = g.pitch;
    b.plane = g.plane;
    b.total = g.total;
    return b;
}

// storage box of a level from the [lo,hi) bounds of its non-EXTERIOR cells
static void makeGeom(Geom &g, const int64_t res[3], const int64_t lo[3], const int64_t hi[3])
{
    for (int a = 0; a < 3; ++a)
    {
	g.res[a] = res[a];
	g.org[a] = 2 * (floorDiv2(lo[a]) - 1);
	const int end = 2 * (ceilDiv2(hi[a]) + 1);
	g.n[a] = end - g.org[a];
    }
    g.pitch = int(divUp(g.n[0], 16) * 16);
    g.plane = int64_t(g.pitch) * g.n[1];
    g.total = g.plane * g.n[2];
    g.chunksPerPlane = int(divUp(g.plane, CHUNK_CELLS));
    g.zBlocks = int(divUp(g.n[2], CHUNK_Z));
}

// clip of the storage box against the host grid [0, hostRes): storage-coordinate range [lo,hi) that exists on the host
// bounds (nullable): expanded-coordinate box [bounds, bounds+3) outside which the host array is not to be read at all
// (the caller's non-EXTERIOR hint: labels are EXTERIOR, weights and values 0 out there by contract)
static bool clipBox(const Geom &g, const int64_t hostRes[3], int lo[3], int hi[3], const int64_t *bounds = nullptr)
{
    bool any = true;
    for (int a = 0; a < 3; ++a)
    {
	int64_t l = std::max<int64_t>(0, -int64_t(g.org[a]));
	int64_t h = std::min<int64_t>(g.n[a], hostRes[a] - g.org[a]);
	if (bounds)
	{
	    l = std::max<int64_t>(l, bounds[a] - g.org[a]);
	    h = std::min<int64_t>(h, bounds[3 + a] - g.org[a]);
	}
	lo[a] = int(l);
	hi[a] = int(h);
	if (hi[a] <= lo[a]) any = false;
    }
    return any;
}

// Host <-> device box transfers.  The caller's arrays are dense EXPANDED grids in pageable (or pinned) memory, of which
// only the cropped box moves: rows are gathered by a few host threads into two pinned 32 MB buffers and shipped as
// contiguous async copies, double-buffered so the gather of slab k overlaps the DMA of slab k-1.
static int ensurePinned(gmg_ctx *ctx)
{
    if (ctx->pin[0]) return GMG_OK;
    ctx->pinCap = size_t(32) << 20;
    for (int i = 0; i < 2; ++i)
    {
	GMG_CUDA(cudaMallocHost(&ctx->pin[i], ctx->pinCap));
	GMG_CUDA(cudaEventCreateWithFlags(&ctx->pinEv[i], cudaEventDisableTiming));
    }
    return GMG_OK;
}

template <typename Fn>
static void parallelRows(int64_t rows, const Fn &fn)
{
    const int nt = int(std::max<int64_t>(1, std::min<int64_t>(std::min(16u, std::max(1u, std::thread::hardware_concurrency())), rows / 64)));
    if (nt <= 1) { fn(0, rows); return; }
    std::vector<std::thread> th;
    const int64_t per = (rows + nt - 1) / nt;
    for (int t = 0; t < nt; ++t)
    {
	const int64_t r0 = t * per, r1 = std::min(rows, r0 + per);
	if (r0 < r1) th.emplace_back([&fn, r0, r1]() { fn(r0, r1); });
    }
    for (auto &t : th) t.join();
}

// page-locked caller memory (cudaHostAlloc / cudaHostRegister / torch pin_memory) is DMA'd in place: one strided 3D copy
// of the cropped box, no host-side gather
static bool isPinnedHost(const void *p)
{
    static const bool off = [] { const char *e = getenv("GMG_NO_DIRECT_DMA"); return e && e[0] == '1'; }();
    if (off) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
template <typename T>
static int copyBox3D(gmg_ctx *ctx, T *staging, T *host, const int64_t hostRes[3], const Geom &g, const int lo[3], const int hi[3], bool toDevice)
{
    const size_t nx = size_t(hi[0] - lo[0]), ny = size_t(hi[1] - lo[1]), nz = size_t(hi[2] - lo[2]);
    const size_t x0 = size_t(g.org[0] + lo[0]), y0 = size_t(g.org[1] + lo[1]), z0 = size_t(g.org[2] + lo[2]);
    cudaMemcpy3DParms p = {};
    const cudaPitchedPtr hp = make_cudaPitchedPtr(host, size_t(hostRes[0]) * sizeof(T), size_t(hostRes[0]) * sizeof(T), size_t(hostRes[1]));
    const cudaPitchedPtr dp = make_cudaPitchedPtr(staging, nx * sizeof(T), nx * sizeof(T), ny);
    const cudaPos hpos = make_cudaPos(x0 * sizeof(T), y0, z0), dpos = make_cudaPos(0, 0, 0);
    p.srcPtr = toDevice ? hp : dp;
    p.srcPos = toDevice ? hpos : dpos;
    p.dstPtr = toDevice ? dp : hp;
    p.dstPos = toDevice ? dpos : hpos;
    p.extent = make_cudaExtent(nx * sizeof(T), ny, nz);
    p.kind = toDevice ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    GMG_CUDA(cudaMemcpy3DAsync(&p, ctx->stream));
    return GMG_OK;
}

// The same for the sub-boxes of `groups` only (clipped to [lo, hi); zShift = local storage plane of g's plane 0): the rest of the
// staging box is not transferred.  Used for vector grids, whose cells outside the groups are 0 on both sides.
template <typename T>
static int copyGroups3D(gmg_ctx *ctx, T *staging, T *host, const int64_t hostRes[3], const Geom &g, const int lo[3], const int hi[3], bool toDevice,
			const std::vector<IoGroup> &groups, int zShift)
{
    const size_t nx = size_t(hi[0] - lo[0]), ny = size_t(hi[1] - lo[1]);
    const cudaPitchedPtr hp = make_cudaPitchedPtr(host, size_t(hostRes[0]) * sizeof(T), size_t(hostRes[0]) * sizeof(T), size_t(hostRes[1]));
    const cudaPitchedPtr dp = make_cudaPitchedPtr(staging, nx * sizeof(T), nx * sizeof(T), ny);
    for (const IoGroup &q : groups)
    {
	const int a0[3] = {std::max(q.x0, lo[0]), std::max(q.y0, lo[1]), std::max(q.z0 - zShift, lo[2])};
	const int a1[3] = {std::min(q.x1, hi[0]), std::min(q.y1, hi[1]), std::min(q.z1 - zShift, hi[2])};
	if (a1[0] <= a0[0] || a1[1] <= a0[1] || a1[2] <= a0[2]) continue;
	cudaMemcpy3DParms p = {};
	const cudaPos hpos = make_cudaPos(size_t(g.org[0] + a0[0]) * sizeof(T), size_t(g.org[1] + a0[1]), size_t(g.org[2] + a0[2]));
	const cudaPos dpos = make_cudaPos(size_t(a0[0] - lo[0]) * sizeof(T), size_t(a0[1] - lo[1]), size_t(a0[2] - lo[2]));
	p.srcPtr = toDevice ? hp : dp;
	p.srcPos = toDevice ? hpos : dpos;
	p.dstPtr = toDevice ? dp : hp;
	p.dstPos = toDevice ? dpos : hpos;
	p.extent = make_cudaExtent(size_t(a1[0] - a0[0]) * sizeof(T), size_t(a1[1] - a0[1]), size_t(a1[2] - a0[2]));
	p.kind = toDevice ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
	GMG_CUDA(cudaMemcpy3DAsync(&p, ctx->stream));
    }
    return GMG_OK;
}

template <typename T>
static int copyBoxH2D(gmg_ctx *ctx, T *staging, const T *host, const int64_t hostRes[3], const Geom &g, const int lo[3], const int hi[3],
		      const std::vector<IoGroup> *groups = nullptr, int zShift = 0)
{
    if (isPinnedHost(host))
    {
	if (groups && !groups->empty()) return copyGroups3D(ctx, staging, const_cast<T *>(host), hostRes, g, lo, hi, true, *groups, zShift);
	return copyBox3D(ctx, staging, const_cast<T *>(host), hostRes, g, lo, hi, true);
    }
    GMG_TRY(ensurePinned(ctx));
    const int64_t nx = hi[0] - lo[0], ny = hi[1] - lo[1], nz = hi[2] - lo[2];
    const int64_t x0 = g.org[0] + lo[0], y0 = g.org[1] + lo[1], z0 = g.org[2] + lo[2];
    const int64_t slabBytes = nx * ny * int64_t(sizeof(T));
    const int64_t zPer = std::max<int64_t>(1, int64_t(ctx->pinCap) / std::max<int64_t>(slabBytes, 1));
    if (slabBytes > int64_t(ctx->pinCap)) return gmg::cudaFail(cudaErrorInvalidValue, "box slab larger than the pinned staging buffer", __FILE__, __LINE__);
    int buf = 0;
    for (int64_t zs = 0; zs < nz; zs += zPer, buf ^= 1)
    {
	const int64_t zc = std::min(zPer, nz - zs);
	GMG_CUDA(cudaEventSynchronize(ctx->pinEv[buf]));
	T *pin = static_cast<T *>(ctx->pin[buf]);
	parallelRows(zc * ny, [&](int64_t r0, int64_t r1) {
	    for (int64_t r = r0; r < r1; ++r)
	    {
		const int64_t z = r / ny, y = r - z * ny;
		std::memcpy(pin + r * nx, host + x0 + hostRes[0] * ((y0 + y) + hostRes[1] * (z0 + zs + z)), size_t(nx) * sizeof(T));
	    }
	});
	GMG_CUDA(cudaMemcpyAsync(staging + zs * nx * ny, pin, size_t(zc * slabBytes), cudaMemcpyHostToDevice, ctx->stream));
	GMG_CUDA(cudaEventRecord(ctx->pinEv[buf], ctx->stream));
    }
    return GMG_OK;
}
template <typename T>
static int copyBoxD2H(gmg_ctx *ctx, T *host, const T *staging, const int64_t hostRes[3], const Geom &g, const int lo[3], const int hi[3],
		      const std::vector<IoGroup> *groups = nullptr, int zShift = 0)
{
    if (isPinnedHost(host))
    {
	if (groups && !groups->empty()) return copyGroups3D(ctx, const_cast<T *>(staging), host, hostRes, g, lo, hi, false, *groups, zShift);
	return copyBox3D(ctx, const_cast<T *>(staging), host, hostRes, g, lo, hi, false);
    }
    GMG_TRY(ensurePinned(ctx));
    const int64_t nx = hi[0] - lo[0], ny = hi[1] - lo[1], nz = hi[2] - lo[2];
    const int64_t x0 = g.org[0] + lo[0], y0 = g.org[1] + lo[1], z0 = g.org[2] + lo[2];
    const int64_t slabBytes = nx * ny * int64_t(sizeof(T));
    const int64_t zPer = std::max<int64_t>(1, int64_t(ctx->pinCap) / std::max<int64_t>(slabBytes, 1));
    if (slabBytes > int64_t(ctx->pinCap)) return gmg::cudaFail(cudaErrorInvalidValue, "box slab larger than the pinned staging buffer", __FILE__, __LINE__);
    // issue slab k+1's DMA before scattering slab k
    auto issue = [&](int64_t zs, int buf) -> int {
	const int64_t zc = std::min(zPer, nz - zs);
	GMG_CUDA(cudaMemcpyAsync(ctx->pin[buf], staging + zs * nx * ny, size_t(zc * slabBytes), cudaMemcpyDeviceToHost, ctx->stream));
	GMG_CUDA(cudaEventRecord(ctx->pinEv[buf], ctx->stream));
	return GMG_OK;
    };
    int buf = 0;
    if (nz > 0) GMG_TRY(issue(0, 0));
    for (int64_t zs = 0; zs < nz; zs += zPer, buf ^= 1)
    {
	const int64_t zc = std::min(zPer, nz - zs);
	if (zs + zPer < nz) GMG_TRY(issue(zs + zPer, buf ^ 1));
	GMG_CUDA(cudaEventSynchronize(ctx->pinEv[buf]));
	const T *pin = static_cast<const T *>(ctx->pin[buf]);
	parallelRows(zc * ny, [&](int64_t r0, int64_t r1) {
	    for (int64_t r = r0; r < r1; ++r)
	    {
		const int64_t z = r / ny, y = r - z * ny;
		std::memcpy(host + x0 + hostRes[0] * ((y0 + y) + hostRes[1] * (z0 + zs + z)), pin + r * nx, size_t(nx) * sizeof(T));
	    }
	});
    }
    return GMG_OK;
}

// host scan for the [lo,hi) bounds of non-EXTERIOR labels (used when the caller gives no hint)
template <typename LabelT>
static bool scanBounds(const LabelT *labels, const int64_t res[3], int64_t lo[3], int64_t hi[3])
{
    const int nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::array<int64_t, 6>> part(nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
	th.emplace_back([&, t]() {
	    std::array<int64_t, 6> b = {res[0], res[1], res[2], -1, -1, -1};
	    for (int64_t z = t; z < res[2]; z += nt)
		for (int64_t y = 0; y < res[1]; ++y)
		{
		    const LabelT *row = labels + res[0] * (y + res[1] * z);
		    int64_t x0 = 0, x1 = res[0] - 1;
		    while (x0 <= x1 && row[x0] == L_EXTERIOR) ++x0;
		    if (x0 > x1) continue;
		    while (row[x1] == L_EXTERIOR) --x1;
		    b[0] = std::min(b[0], x0); b[3] = std::max(b[3], x1);
		    b[1] = std::min(b[1], y); b[4] = std::max(b[4], y);
		    b[2] = std::min(b[2], z); b[5] = std::max(b[5], z);
		}
	    part[t] = b;
	});
    for (auto &t : th) t.join();
    std::array<int64_t, 6> b = {res[0], res[1], res[2], -1, -1, -1};
    for (auto &p : part)
	for (int a = 0; a < 3; ++a) { b[a] = std::min(b[a], p[a]); b[a + 3] = std::max(b[a + 3], p[a + 3]); }
    if (b[3] < 0) return false;
    for (int a = 0; a < 3; ++a) { lo[a] = b[a]; hi[a] = b[a + 3] + 1; }
    return true;
}

template <typename LabelT>
static int uploadLabels(gmg_ctx *ctx, uint8_t *dst, const LabelT *host, const int64_t res[3], const Geom &g, const int64_t *bounds = nullptr)
{
    int lo[3], hi[3];
    const BoxArgs ba = boxArgs(g);
    if (!clipBox(g, res, lo, hi, bounds))
    {
	GMG_CUDA(cudaMemsetAsync(dst, L_EXTERIOR, g.total, ctx->stream));
	return GMG_OK;
    }
    LabelT *staging = nullptr;
    const int64_t cnt = int64_t(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    GMG_CUDA(devMalloc(&staging, sizeof(LabelT) * cnt));
    GMG_TRY(copyBoxH2D(ctx, staging, host, res, g, lo, hi));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_labels_from_host<LabelT><<<unsigned(divUp(g.total, BLOCK)), BLOCK, 0, ctx->stream>>>(dst, staging, ba, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    }
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(staging));
    return GMG_OK;
}

static int downloadLabels(gmg_ctx *ctx, int32_t *host, const uint8_t *src, const int64_t res[3], const Geom &g, bool fillExterior,
			  const int64_t *bounds = nullptr)
{
    if (fillExterior)
    {
	const int64_t n = res[0] * res[1] * res[2];
	std::fill(host, host + n, int32_t(L_EXTERIOR));
    }
    int lo[3], hi[3];
    if (!clipBox(g, res, lo, hi, bounds)) return GMG_OK;
    int32_t *staging = nullptr;
    const int64_t cnt = int64_t(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    GMG_CUDA(devMalloc(&staging, sizeof(int32_t) * cnt));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_labels_to_i32<<<unsigned(divUp(cnt, BLOCK)), BLOCK, 0, ctx->stream>>>(staging, src, boxArgs(g), lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    }
    GMG_TRY(copyBoxD2H(ctx, host, staging, res, g, lo, hi));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(staging));
    return GMG_OK;
}

// host dense grid (hostRes may be a face grid) -> pitched storage, zero where the host has no value
// groups (only together with maskLabels, which zeroes whatever the untransferred part of the staging box holds): move the
// active rectangles only
static int uploadValues(gmg_ctx *ctx, double *dst, const double *host, const int64_t hostRes[3], const Geom &g, const uint8_t *maskLabels = nullptr,
			const int64_t *bounds = nullptr, const std::vector<IoGroup> *groups = nullptr, int zShift = 0)
{
    if (!maskLabels) groups = nullptr;
    int lo[3], hi[3];
    if (!clipBox(g, hostRes, lo, hi, bounds))
    {
	GMG_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * g.total, ctx->stream));
	return GMG_OK;
    }
    double *staging = nullptr;
    const int64_t cnt = int64_t(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    GMG_CUDA(devMalloc(&staging, sizeof(double) * cnt));
    GMG_TRY(copyBoxH2D(ctx, staging, host, hostRes, g, lo, hi, groups, zShift));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_values_from_staging<<<unsigned(divUp(g.total, BLOCK)), BLOCK, 0, ctx->stream>>>(dst, staging, maskLabels, boxArgs(g), lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    }
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(staging));
    return GMG_OK;
}

static int downloadValues(gmg_ctx *ctx, double *host, const double *src, const int64_t hostRes[3], const Geom &g, bool fillZero,
			  const int64_t *bounds = nullptr, const std::vector<IoGroup> *groups = nullptr, int zShift = 0)
{
    if (fillZero) std::memset(host, 0, sizeof(double) * size_t(hostRes[0]) * hostRes[1] * hostRes[2]);
    int lo[3], hi[3];
    if (!clipBox(g, hostRes, lo, hi, bounds)) return GMG_OK;
    double *staging = nullptr;
    const int64_t cnt = int64_t(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    GMG_CUDA(devMalloc(&staging, sizeof(double) * cnt));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_values_to_staging<<<unsigned(divUp(cnt, BLOCK)), BLOCK, 0, ctx->stream>>>(staging, src, boxArgs(g), lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    }
    GMG_TRY(copyBoxD2H(ctx, host, staging, hostRes, g, lo, hi, groups, zShift));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(staging));
    return GMG_OK;
}

// ====================================================================================================
// device-side builders shared by the solver constructor and the stand-alone builder entry points
// ====================================================================================================
static int countLabels(gmg_ctx *ctx, const uint8_t *labels, int64_t total, int64_t *nInterior, int64_t *nBoundary)
{
    unsigned long long *d = nullptr, h[2] = {0, 0};
    GMG_CUDA(devMalloc(&d, 2 * sizeof(unsigned long long)));
    GMG_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), ctx->stream));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	const unsigned grid = unsigned(std::min<int64_t>(divUp(total, BLOCK), 4096));
	k_count_labels<<<grid, BLOCK, 0, ctx->stream>>>(labels, total, d);
    }
    GMG_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(d));
    *nInterior = int64_t(h[0]);
    *nBoundary = int64_t(h[1]);
    return GMG_OK;
}

// compaction of the indices i in [0,n) with flags[i] != 0 (ascending)
static int selectFlagged(gmg_ctx *ctx, const uint8_t *flags, int64_t n, int32_t **out, int *count)
{
    int *dCount = nullptr;
    int32_t *tmpOut = nullptr;
    GMG_CUDA(devMalloc(&dCount, sizeof(int)));
    GMG_CUDA(devMalloc(&tmpOut, sizeof(int32_t) * std::max<int64_t>(n, 1)));
    void *dTemp = nullptr;
    size_t tempBytes = 0;
    thrust::counting_iterator<int32_t> it(0);
    GMG_CUDA(cub::DeviceSelect::Flagged(dTemp, tempBytes, it, flags, tmpOut, dCount, int(n), ctx->stream));
    GMG_CUDA(devMalloc(&dTemp, std::max<size_t>(tempBytes, 16)));
    GMG_CUDA(cub::DeviceSelect::Flagged(dTemp, tempBytes, it, flags, tmpOut, dCount, int(n), ctx->stream));
    ++ctx->launches;
    int h = 0;
    GMG_CUDA(cudaMemcpyAsync(&h, dCount, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = h;
    *out = nullptr;
    GMG_CUDA(devMalloc(out, sizeof(int32_t) * std::max(h, 1)));
    GMG_CUDA(cudaMemcpyAsync(*out, tmpOut, sizeof(int32_t) * h, cudaMemcpyDeviceToDevice, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(dTemp));
    GMG_CUDA(devFree(tmpOut));
    GMG_CUDA(devFree(dCount));
    return GMG_OK;
}

// boundary band of one level: BOUNDARY cells first, then the INTERIOR cells within width-1 steps (Ops.cpp:165-469)
static int buildBand(gmg_ctx *ctx, Level &L, int width)
{
    // the band mask is grown over the level's GLOBAL box (a slab edge must not clip the dilation); the lists are
    // compacted over the rank's stored planes and hold local storage indices
    const Geom &g = L.g;
    const Geom &gg = L.gg;
    const uint8_t *labelsG = L.labelsAlloc ? L.labelsAlloc : L.labels;
    const BoxArgs bg = boxArgs(gg);
    const unsigned gridG = unsigned(divUp(gg.total, BLOCK));
    const unsigned grid = unsigned(divUp(g.total, BLOCK));
    uint8_t *m0 = nullptr, *m1 = nullptr;
    GMG_CUDA(devMalloc(&m0, gg.total));
    GMG_CUDA(devMalloc(&m1, gg.total));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_band_init<<<gridG, BLOCK, 0, ctx->stream>>>(m0, labelsG, gg.total);
    }
    for (int layer = 0; layer < width - 1; ++layer)
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_band_dilate<<<gridG, BLOCK, 0, ctx->stream>>>(m1, m0, labelsG, bg);
	std::swap(m0, m1);
    }
    {
	// the band flags the zero-aware interior sweep reads (SM_JACOBI_ZERO), over the global box like the labels
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_band_dilate<<<gridG, BLOCK, 0, ctx->stream>>>(m1, m0, labelsG, bg);
	GMG_CUDA(devMalloc(&L.flagsAlloc, gg.total));
	k_band_flag_grid<<<gridG, BLOCK, 0, ctx->stream>>>(L.flagsAlloc, m0, m1, gg.total);
	L.bandFlags = L.flagsAlloc + int64_t(L.zOff) * g.plane;
	GMG_CUDA(devMalloc(&L.nbrMaskAlloc, gg.total));
	k_band_nbr_mask<<<gridG, BLOCK, 0, ctx->stream>>>(L.nbrMaskAlloc, m0, labelsG, bg);
	L.nbrMask = L.nbrMaskAlloc + int64_t(L.zOff) * g.plane;
    }
    const uint8_t *mLocal = m0 + int64_t(L.zOff) * g.plane;
    int32_t *idxB = nullptr, *idxI = nullptr;
    int nB = 0, nI = 0;
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_band_flags<<<grid, BLOCK, 0, ctx->stream>>>(m1, mLocal, L.labels, 0, g.total);
    }
    GMG_TRY(selectFlagged(ctx, m1, g.total, &idxB, &nB));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_band_flags<<<grid, BLOCK, 0, ctx->stream>>>(m1, mLocal, L.labels, 1, g.total);
    }
    GMG_TRY(selectFlagged(ctx, m1, g.total, &idxI, &nI));
    GMG_CUDA(devFree(m0));
    GMG_CUDA(devFree(m1));
    L.nBoundary = nB;
    L.nBand = nB + nI;
    const int nBand = L.nBand;
    // ONE slab per level for everything the band sweeps touch (indices, neighbour references, the three compact value
    // arrays, the coefficient records): a single address range that can be given an L2 persisting access-policy window
    // (bandWindow below), so the twelve sweeps of a V-cycle find their metadata in L2 instead of re-reading it from HBM
    // after every full-grid pass has streamed through the cache.
    {
	size_t off = 0;
	auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
	const size_t n1 = size_t(std::max(nBand, 1)), nb1 = size_t(std::max(nB, 1));
	const size_t oIdx = take(sizeof(int32_t) * n1), oRef = take(sizeof(int32_t) * 6 * n1), oV0 = take(sizeof(double) * n1), oV1 = take(sizeof(double) * n1),
		     oB = take(sizeof(double) * n1), oCoef = take(sizeof(double) * 8 * nb1), oCode = take(sizeof(unsigned short) * nb1);
	char *slab = nullptr;
	GMG_CUDA(devMalloc(&slab, off));
	L.bandSlab = slab;
	L.bandSlabBytes = off;
	L.bandIdx = reinterpret_cast<int32_t *>(slab + oIdx);
	L.bandRef = reinterpret_cast<int32_t *>(slab + oRef);
	L.bandV0 = reinterpret_cast<double *>(slab + oV0);
	L.bandV1 = reinterpret_cast<double *>(slab + oV1);
	L.bandB = reinterpret_cast<double *>(slab + oB);
	L.bcoef = reinterpret_cast<double *>(slab + oCoef);
	L.wcode = reinterpret_cast<unsigned short *>(slab + oCode);
    }
    GMG_CUDA(cudaMemcpyAsync(L.bandIdx, idxB, sizeof(int32_t) * nB, cudaMemcpyDeviceToDevice, ctx->stream));
    GMG_CUDA(cudaMemcpyAsync(L.bandIdx + nB, idxI, sizeof(int32_t) * nI, cudaMemcpyDeviceToDevice, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(idxB));
    GMG_CUDA(devFree(idxI));
    if (nBand > 0)
    {
	// one guard plane each side: band cells of a slab's first / last stored plane look one plane out
	int32_t *posBase = nullptr;
	const int64_t posN = g.total + 2 * g.plane;
	GMG_CUDA(devMalloc(&posBase, sizeof(int32_t) * posN));
	int32_t *pos = posBase + g.plane;
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_fill_i32<<<unsigned(divUp(posN, BLOCK)), BLOCK, 0, ctx->stream>>>(posBase, -1, posN);
	}
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_band_pos<<<unsigned(divUp(nBand, BLOCK)), BLOCK, 0, ctx->stream>>>(pos, L.bandIdx, nBand);
	}
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_band_ref<<<unsigned(divUp(nBand, BLOCK)), BLOCK, 0, ctx->stream>>>(L.bandRef, pos, L.bandIdx, L.labels, nBand, g.pitch, g.plane);
	}
	GMG_CUDA(cudaStreamSynchronize(ctx->stream));
	GMG_CUDA(devFree(posBase));
    }
    return GMG_OK;
}

static int buildCoefs(gmg_ctx *ctx, Level &L, const double *w0, const double *w1, const double *w2)
{
    if (L.nBoundary > 0)
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_band_coef<<<unsigned(divUp(L.nBoundary, BLOCK)), BLOCK, 0, ctx->stream>>>(L.bcoef, L.wcode, L.bandIdx, L.nBoundary, L.labels, w0, w1, w2, L.g.pitch,
										   L.g.plane);
    }
    L.hasWeights = w0 != nullptr;
    return GMG_OK;
}

// Coefficient records of level 0 from the caller's HOST weight grids without uploading them: the BOUNDARY cell list comes
// back from the device (4 bytes per cell), the host gathers the six face weights of each cell, and only those travel
// (48 bytes per BOUNDARY cell instead of 24 bytes per cell of the box).  Face weights outside `bounds` (+1 on the face
// axis) read as 0, exactly as the clipped upload of the full grids did.
struct FaceLimits
{
    int64_t fr[3][3];      // extents of the three face-weight grids
    int64_t lim[3][3][2];  // readable range per grid and axis
};

// hw[n][k] (k < nb) = weight of face n of the BOUNDARY cell with storage index idx[k]; 0 outside the readable range.
// The six faces of a cell are six cache / TLB misses in GB-sized arrays: addresses of a block of cells first (with prefetches
// in flight), values afterwards, over a few host threads.
static void gatherFaceWeights(const int32_t *idx, double *hw, int64_t nb, Geom g, const double *w0, const double *w1, const double *w2, FaceLimits fl)
{
    const double *w[3] = {w0, w1, w2};
    parallelRows(nb, [&](int64_t r0, int64_t r1) {
	constexpr int PB = 64;
	const double *addr[PB][6];
	for (int64_t kb = r0; kb < r1; kb += PB)
	{
	    const int m = int(std::min<int64_t>(PB, r1 - kb));
	    for (int q = 0; q < m; ++q)
	    {
		const int64_t i = idx[kb + q];
		const int64_t z = i / g.plane, rem = i - z * g.plane, y = rem / g.pitch, x = rem - y * g.pitch;
		const int64_t e[3] = {x + g.org[0], y + g.org[1], z + g.org[2]};
		for (int n = 0; n < 6; ++n)
		{
		    const int a = n >> 1;
		    int64_t f[3] = {e[0], e[1], e[2]};
		    if (n & 1) ++f[a];
		    const bool inside = f[0] >= fl.lim[a][0][0] && f[0] < fl.lim[a][0][1] && f[1] >= fl.lim[a][1][0] && f[1] < fl.lim[a][1][1] &&
					f[2] >= fl.lim[a][2][0] && f[2] < fl.lim[a][2][1];
		    addr[q][n] = inside ? w[a] + (f[2] * fl.fr[a][1] + f[1]) * fl.fr[a][0] + f[0] : nullptr;
		    if (inside) __builtin_prefetch(addr[q][n], 0, 0);
		}
	    }
	    for (int q = 0; q < m; ++q)
		for (int n = 0; n < 6; ++n) hw[int64_t(n) * nb + kb + q] = addr[q][n] ? *addr[q][n] : 0.0;
	}
    });
}

static FaceLimits faceLimits(const int64_t res[3], const int64_t *bounds)
{
    FaceLimits fl;
    for (int a = 0; a < 3; ++a)
	for (int c = 0; c < 3; ++c)
	{
	    fl.fr[a][c] = res[c] + (c == a ? 1 : 0);
	    fl.lim[a][c][0] = bounds ? bounds[c] : 0;
	    fl.lim[a][c][1] = std::min<int64_t>((bounds ? bounds[3 + c] : res[c]) + (c == a ? 1 : 0), fl.fr[a][c]);
	}
    return fl;
}

// Pure host function behind the sparse face weights (no device needed): out[n * count + k] = weight of face n
// (-x,+x,-y,+y,-z,+z) of the cell with storage index idx[k] in a box of row pitch `pitch`, plane size `plane` and expanded
// origin org; 0 outside bounds (lo[3], hi[3]; null = the whole grids).  w0/w1/w2 have one more entry along their own axis.
extern "C" int gmg_gather_face_weights(const int32_t *idx, int64_t count, int pitch, int64_t plane, const int32_t org[3], const double *w0,
				       const double *w1, const double *w2, const int64_t res[3], const int64_t *bounds, double *out)
{
    if (!idx || !org || !w0 || !w1 || !w2 || !res || !out || count < 0 || pitch <= 0 || plane <= 0) return invalid("gmg_gather_face_weights: bad argument");
    Geom g = {};
    g.pitch = pitch;
    g.plane = plane;
    for (int a = 0; a < 3; ++a) g.org[a] = org[a];
    gatherFaceWeights(idx, out, count, g, w0, w1, w2, faceLimits(res, bounds));
    return GMG_OK;
}

// host gather of the level-0 face weights running beside the rest of the setup (joined by finishCoefsSparse)
struct CoefJob
{
    std::thread th;
    bool pending = false;
    ~CoefJob() { if (th.joinable()) th.join(); }
};

static int coefKernelSparse(gmg_ctx *ctx, Level &L, const double *dw)
{
    GMG_LAUNCH(ctx, KC_SETUP, 0);
    k_band_coef_sparse<<<unsigned(divUp(L.nBoundary, BLOCK)), BLOCK, 0, ctx->stream>>>(L.bcoef, L.wcode, L.bandIdx, L.nBoundary, L.labels, dw, L.g.pitch, L.g.plane);
    return GMG_OK;
}

static int buildCoefsSparse(gmg_ctx *ctx, Level &L, const double *w0, const double *w1, const double *w2, const int64_t res[3], const int64_t *bounds,
			    CoefJob *job)
{
    const int nB = L.nBoundary;
    L.hasWeights = true;
    if (nB == 0) return GMG_OK;
    const Geom &g = L.g;
    GMG_TRY(ensurePinned(ctx));
    const FaceLimits fl = faceLimits(res, bounds);
    // pinned buffer 0 receives the cell indices of a batch, pinned buffer 1 carries its face weights
    const int64_t batch = std::min<int64_t>(int64_t(ctx->pinCap / sizeof(int32_t)), int64_t(ctx->pinCap / (6 * sizeof(double))));
    int32_t *idx = static_cast<int32_t *>(ctx->pin[0]);
    double *hw = static_cast<double *>(ctx->pin[1]);
    if (job && nB <= batch)
    {
	// one batch: the gather runs on its own thread while the constructor builds the other levels
	GMG_CUDA(cudaMemcpyAsync(idx, L.bandIdx, sizeof(int32_t) * nB, cudaMemcpyDeviceToHost, ctx->stream));
	GMG_CUDA(cudaStreamSynchronize(ctx->stream));
	const Geom gc = g;
	job->th = std::thread([=]() {
	    const double t0 = nowMs();
	    gatherFaceWeights(idx, hw, nB, gc, w0, w1, w2, fl);
	    if (getenv("GMG_PRINT_STATS")) printf("      [face-weight gather thread: %d cells, %.3f ms]\n", nB, nowMs() - t0);
	});
	job->pending = true;
	return GMG_OK;
    }
    double *dw = nullptr;
    GMG_CUDA(devMalloc(&dw, sizeof(double) * 6 * nB));
    for (int64_t k0 = 0; k0 < nB; k0 += batch)
    {
	const int64_t nb = std::min<int64_t>(batch, nB - k0);
	GMG_CUDA(cudaMemcpyAsync(idx, L.bandIdx + k0, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, ctx->stream));
	GMG_CUDA(cudaStreamSynchronize(ctx->stream));
	gatherFaceWeights(idx, hw, nb, g, w0, w1, w2, fl);
	for (int n = 0; n < 6; ++n)
	    GMG_CUDA(cudaMemcpyAsync(dw + int64_t(n) * nB + k0, hw + int64_t(n) * nb, sizeof(double) * nb, cudaMemcpyHostToDevice, ctx->stream));
	GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    GMG_TRY(coefKernelSparse(ctx, L, dw));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(dw));
    return GMG_OK;
}

static int finishCoefsSparse(gmg_ctx *ctx, Level &L, CoefJob *job)
{
    if (!job || !job->pending) return GMG_OK;
    job->th.join();
    job->pending = false;
    // the kernel reads the gathered weights straight out of the pinned buffer (mapped under unified addressing; coalesced):
    // measured 0.4 ms for this step at 256^3, against 1.6 ms with an explicit 10 MB copy of the just-written buffer first
    // (profiles/r01_setup_phases.txt)
    GMG_TRY(coefKernelSparse(ctx, L, static_cast<const double *>(ctx->pin[1])));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMG_OK;
}

// Pure host function behind the transfer plan (no device needed): extents[z] = (x0, x1, y0, y1) of plane z's active cells,
// half-open, x1 <= x0 for an empty plane.  Consecutive non-empty planes are merged into one box while the merged box holds at
// most 15 % more cells than the planes' own rectangles; groups[k] = (z0, z1, x0, x1, y0, y1).
extern "C" int gmg_transfer_plan(const int32_t *extents, int planes, int32_t *groups, int *groupCount, int64_t *cells)
{
    if (!extents || !groups || !groupCount || planes < 0) return invalid("gmg_transfer_plan: bad argument");
    int n = 0;
    int64_t total = 0;
    IoGroup cur = {0, 0, 0, 0, 0, 0};
    int64_t own = 0;  // sum of the member planes' own rectangle areas
    auto flush = [&]() {
	if (cur.z1 > cur.z0)
	{
	    int32_t *q = groups + size_t(6) * n++;
	    q[0] = cur.z0; q[1] = cur.z1; q[2] = cur.x0; q[3] = cur.x1; q[4] = cur.y0; q[5] = cur.y1;
	    total += int64_t(cur.x1 - cur.x0) * (cur.y1 - cur.y0) * (cur.z1 - cur.z0);
	}
	cur.z0 = cur.z1 = 0;
	own = 0;
    };
    for (int z = 0; z < planes; ++z)
    {
	const int32_t *r = extents + size_t(4) * z;
	if (r[1] <= r[0] || r[3] <= r[2]) { flush(); continue; }
	const int64_t area = int64_t(r[1] - r[0]) * (r[3] - r[2]);
	if (cur.z1 > cur.z0)
	{
	    IoGroup m = cur;
	    m.x0 = std::min(m.x0, r[0]); m.x1 = std::max(m.x1, r[1]); m.y0 = std::min(m.y0, r[2]); m.y1 = std::max(m.y1, r[3]); m.z1 = z + 1;
	    const int64_t vol = int64_t(m.x1 - m.x0) * (m.y1 - m.y0) * (m.z1 - m.z0);
	    if (double(vol) <= 1.15 * double(own + area)) { cur = m; own += area; continue; }
	    flush();
	}
	cur = {z, z + 1, r[0], r[1], r[2], r[3]};
	own = area;
    }
    flush();
    *groupCount = n;
    if (cells) *cells = total;
    return GMG_OK;
}

// Transfer plan of the level-0 vector grids (rhs, initial guess, pressure): per z-plane the bounding rectangle of the active
// cells, consecutive planes merged into one 3D copy while the merged box stays within 15 % of the planes' own rectangles.
static int buildIoGroups(gmg_solver *s)
{
    s->ioGroups.clear();
    s->ioCells = 0;
    static const bool off = [] { const char *e = getenv("GMG_IO_GROUPS"); return e && e[0] == '0'; }();
    gmg_ctx *ctx = s->ctx;
    const Level &L = s->lv[0];
    const Geom &g = L.g;
    const int nz = g.n[2];
    if (off || nz == 0) { s->ioCells = g.total; return GMG_OK; }
    int4 *d = nullptr;
    GMG_CUDA(devMalloc(&d, sizeof(int4) * nz));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_plane_extents<<<nz, BLOCK, 0, ctx->stream>>>(d, L.labels, g.n[0], g.n[1], g.pitch, g.plane);
    }
    std::vector<int4> e(nz);
    GMG_CUDA(cudaMemcpyAsync(e.data(), d, sizeof(int4) * nz, cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(d));
    std::vector<int32_t> groups(size_t(6) * nz);
    int nGroups = 0;
    GMG_TRY(gmg_transfer_plan(reinterpret_cast<const int32_t *>(e.data()), nz, groups.data(), &nGroups, &s->ioCells));
    for (int k = 0; k < nGroups; ++k)
    {
	const int32_t *q = groups.data() + size_t(6) * k;
	s->ioGroups.push_back({q[0], q[1], q[2], q[3], q[4], q[5]});
    }
    return GMG_OK;
}

static int buildChunks(gmg_ctx *ctx, Level &L)
{
    const Geom &g = L.g;
    const int nChunks = g.chunksPerPlane * g.zBlocks;
    uint8_t *fi = nullptr, *fa = nullptr;
    GMG_CUDA(devMalloc(&fi, nChunks));
    GMG_CUDA(devMalloc(&fa, nChunks));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_chunk_flags<<<nChunks, BLOCK, 0, ctx->stream>>>(fi, fa, L.labels, g.chunksPerPlane, g.plane, g.n[2]);
    }
    GMG_TRY(selectFlagged(ctx, fi, nChunks, &L.chunksInterior, &L.nChunksInterior));
    GMG_TRY(selectFlagged(ctx, fa, nChunks, &L.chunksActive, &L.nChunksActive));
    GMG_CUDA(devFree(fi));
    GMG_CUDA(devFree(fa));
    return GMG_OK;
}

// tile lists of the tiled Gauss-Seidel smoother (Ops.h:441-448: tiles of the expanded grid, parity of tx+ty+tz)
// TMA path of the full-grid stencil kernels (gmg_kernels.cuh: k_stencil_tma): list of the 64 x 8 x 4 bricks holding an
// INTERIOR cell.  Built for levels of at least GMG_TMA_MIN_CELLS cells when GMG_TMA=1 (A/B switch, profiles/).
static int tmaMode()
{
    // read at solver creation; a bit mask of the kernels that take the TMA-staged variant on big levels: 1 interior Jacobi,
    // 2 residual, 4 apply, 8 restriction, 16 prolongation; GMG_TMA=0 forces the plain-load kernels everywhere.
    // Default 23 = all but the restriction.  Measured 512^3 V-cycle, 30 steps (profiles/r02_tma_ab.md): none 5.30 ms, Jacobi
    // 5.22, residual 5.27, prolongation 5.20, those three 5.09; the TMA restriction alone 6.04 -- shared-memory-bound on its
    // 48 loads per coarse cell (ncu: mio_throttle), no faster than k_restrict by itself and slower in the pipeline.
    const char *e = getenv("GMG_TMA");
    return e ? atoi(e) : 23;
}
static int buildBricks(gmg_ctx *ctx, Level &L)
{
    // default: levels whose grids do not fit L2 (measured, profiles/r02_tma_ab.md: residual 0.78 -> 0.89 of the HBM peak at
    // 512^3; at 256^3, where a level-0 grid is 34 MB and L2-resident, the brick kernel's longer prologue loses 15 %)
    const char *mc = getenv("GMG_TMA_MIN_CELLS");
    const int64_t minCells = mc ? atoll(mc) : int64_t(12) << 20;
    // (the level's GLOBAL cell count decides, so a z-slab of a sharded level takes the same kernels as the unsharded level)
    if (!tmaMode() || L.nActiveGlobal < minCells) return GMG_OK;
    const Geom &g = L.g;
    L.bricksX = int(divUp(g.n[0], TB_X));
    L.bricksY = int(divUp(g.n[1], TB_Y));
    const int bz = int(divUp(g.n[2], TB_Z));
    const int64_t nb = int64_t(L.bricksX) * L.bricksY * bz;
    uint8_t *flags = nullptr;
    GMG_CUDA(devMalloc(&flags, size_t(nb)));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_brick_flags<<<unsigned(nb), BLOCK, 0, ctx->stream>>>(flags, L.labels, L.bricksX, L.bricksY, g.pitch, g.plane, g.n[1], g.n[2]);
    }
    GMG_TRY(selectFlagged(ctx, flags, nb, &L.bricks, &L.nBricks));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_brick_flags_active<<<unsigned(nb), BLOCK, 0, ctx->stream>>>(flags, L.labels, L.bricksX, L.bricksY, g.pitch, g.plane, g.n[1], g.n[2]);
    }
    GMG_TRY(selectFlagged(ctx, flags, nb, &L.bricksActive, &L.nBricksActive));
    GMG_CUDA(devFree(flags));
    return GMG_OK;
}
// coarse bricks of level C (32 x 4 x 2 coarse cells holding an active cell) for the TMA restriction INTO it
static int buildCoarseBricks(gmg_ctx *ctx, Level &C)
{
    const Geom &g = C.g;
    C.cbricksX = int(divUp(g.n[0], RB_X));
    C.cbricksY = int(divUp(g.n[1], RB_Y));
    const int bz = int(divUp(g.n[2], RB_Z));
    const int64_t nb = int64_t(C.cbricksX) * C.cbricksY * bz;
    uint8_t *flags = nullptr;
    GMG_CUDA(devMalloc(&flags, size_t(nb)));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_cbrick_flags<<<unsigned(nb), BLOCK, 0, ctx->stream>>>(flags, C.labels, C.cbricksX, C.cbricksY, g.pitch, g.plane, g.n[1], g.n[2]);
    }
    GMG_TRY(selectFlagged(ctx, flags, nb, &C.cbricks, &C.nCBricks));
    GMG_CUDA(devFree(flags));
    return GMG_OK;
}

// 3D tensor map over a vector grid INCLUDING its two guard planes (the tensor's plane 0 is the lower guard plane), box =
// brick + halo.  Cached per grid pointer: the kernels take it by value (__grid_constant__), so a captured graph keeps its own copy.
static int tensorMapOf(gmg_solver *s, int level, const double *grid, TmaMap *out, int boxKind = 0)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
				 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
	void *fn = nullptr;
	cudaDriverEntryPointQueryResult q;
	if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); fn = nullptr; }
	return reinterpret_cast<EncodeFn>(fn);
    }();
    if (!encode) return invalid("cuTensorMapEncodeTiled is not available in this driver");
    auto key = std::make_pair(level * 4 + boxKind, grid);
    auto it = s->tensorMaps.find(key);
    if (it == s->tensorMaps.end())
    {
	const Geom &g = s->lv[level].g;
	static_assert(sizeof(CUtensorMap) == sizeof(TmaMap), "TmaMap must mirror CUtensorMap");
	TmaMap m;
	const cuuint64_t dims[3] = {cuuint64_t(g.pitch), cuuint64_t(g.n[1]), cuuint64_t(g.n[2] + 2)};
	const cuuint64_t strides[2] = {cuuint64_t(g.pitch) * sizeof(double), cuuint64_t(g.plane) * sizeof(double)};
	// boxKind 0: a 64 x 8 x 4 brick plus its stencil / restriction halo; 1: the coarse box a fine brick interpolates from
	const cuuint32_t box[3] = {cuuint32_t(boxKind ? PB_BOX_X : TB_BOX_X), cuuint32_t(boxKind ? PB_BOX_Y : TB_BOX_Y), cuuint32_t(boxKind ? PB_BOX_Z : TB_BOX_Z)};
	const cuuint32_t estr[3] = {1, 1, 1};
	const CUresult r = encode(reinterpret_cast<CUtensorMap *>(&m), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(grid) - g.plane, dims, strides, box, estr,
				  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS)
	{
	    char buf[128];
	    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
	    return invalid(buf);
	}
	it = s->tensorMaps.emplace(key, m).first;
    }
    *out = it->second;
    return GMG_OK;
}

static int buildGsTiles(gmg_ctx *ctx, Level &L)
{
    const Geom &g = L.g;
    auto floorDiv16 = [](int v) { return v >= 0 ? v / 16 : -((-v + 15) / 16); };
    int t0[3], nt[3];
    for (int a = 0; a < 3; ++a)
    {
	t0[a] = floorDiv16(g.org[a]);
	nt[a] = int(divUp(g.org[a] + g.n[a], 16)) - t0[a];
	L.gsOff[a] = t0[a] * 16 - g.org[a];
    }
    L.gsTilesX = nt[0];
    L.gsTilesY = nt[1];
    const int nTiles = nt[0] * nt[1] * nt[2];
    uint8_t *fo = nullptr, *fe = nullptr;
    GMG_CUDA(devMalloc(&fo, nTiles));
    GMG_CUDA(devMalloc(&fe, nTiles));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_gs_tile_flags<<<nTiles, BLOCK, 0, ctx->stream>>>(fo, fe, L.labels, nt[0], nt[1], L.gsOff[0], L.gsOff[1], L.gsOff[2], g.n[0], g.n[1], g.n[2], g.pitch,
							  g.plane, (t0[0] + t0[1] + t0[2]) & 1);
    }
    GMG_TRY(selectFlagged(ctx, fe, nTiles, &L.gsTiles[0], &L.nGsTiles[0]));
    GMG_TRY(selectFlagged(ctx, fo, nTiles, &L.gsTiles[1], &L.nGsTiles[1]));
    GMG_CUDA(devFree(fo));
    GMG_CUDA(devFree(fe));
    GMG_CUDA(devMalloc(&L.bpos, sizeof(int32_t) * g.total));
    if (L.nBoundary > 0)
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_band_pos<<<unsigned(divUp(L.nBoundary, BLOCK)), BLOCK, 0, ctx->stream>>>(L.bpos, L.bandIdx, L.nBoundary);
    }
    return GMG_OK;
}

// band list in the reference's order (tile, z, y, x) and expanded coordinates
// (a sharded level exports the cells of the rank's owned planes; the ranks' lists tile the global one)
static int exportBand(gmg_ctx *ctx, const Level &L, int64_t *xyz, int64_t *count)
{
    const int n = L.nBand;
    if (count) *count = 0;
    if (n == 0) return GMG_OK;
    unsigned long long *keys = nullptr, *keysOut = nullptr;
    GMG_CUDA(devMalloc(&keys, sizeof(unsigned long long) * n));
    GMG_CUDA(devMalloc(&keysOut, sizeof(unsigned long long) * n));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_band_keys<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(keys, L.bandIdx, n, boxArgs(L.g));
    }
    void *dTemp = nullptr;
    size_t tempBytes = 0;
    GMG_CUDA(cub::DeviceRadixSort::SortKeys(dTemp, tempBytes, keys, keysOut, n, 0, 64, ctx->stream));
    GMG_CUDA(devMalloc(&dTemp, std::max<size_t>(tempBytes, 16)));
    GMG_CUDA(cub::DeviceRadixSort::SortKeys(dTemp, tempBytes, keys, keysOut, n, 0, 64, ctx->stream));
    ++ctx->launches;
    std::vector<unsigned long long> h(n);
    GMG_CUDA(cudaMemcpyAsync(h.data(), keysOut, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    const int64_t zLo = int64_t(L.g.org[2]) + L.ownLo, zHi = int64_t(L.g.org[2]) + L.ownHi;
    int64_t m = 0;
    for (int k = 0; k < n; ++k)
    {
	const int64_t ez = int64_t((h[k] >> 24) & 0xfff);
	if (L.sharded && (ez < zLo || ez >= zHi)) continue;
	if (xyz)
	{
	    xyz[3 * m] = int64_t(h[k] & 0xfff);
	    xyz[3 * m + 1] = int64_t((h[k] >> 12) & 0xfff);
	    xyz[3 * m + 2] = ez;
	}
	++m;
    }
    if (count) *count = m;
    GMG_CUDA(devFree(dTemp));
    GMG_CUDA(devFree(keys));
    GMG_CUDA(devFree(keysOut));
    return GMG_OK;
}

static void freeLevel(Level &L)
{
    devFree(L.labelsAlloc ? L.labelsAlloc : L.labels); devFree(L.bandSlab); devFree(L.flagsAlloc); devFree(L.nbrMaskAlloc);
    delete static_cast<ClusterSmoothArgs *>(L.smoothArgs);
    devFree(L.smoothSlab);
    delete static_cast<BandTileArgs *>(L.tileArgs);
    devFree(L.tileSlab);
    devFree(L.chunksInterior); devFree(L.chunksActive); devFree(L.bricks); devFree(L.bricksActive); devFree(L.cbricks);
    devFree(L.gsTiles[0]); devFree(L.gsTiles[1]); devFree(L.bpos);
    freeGrid(L.x, L.g); freeGrid(L.xAlt, L.g); freeGrid(L.b, L.g); freeGrid(L.r, L.g);
    freeGrid32(L.x32, L.g); freeGrid32(L.xAlt32, L.g); freeGrid32(L.b32, L.g); freeGrid32(L.r32, L.g);
    devFree(L.bandV0f); devFree(L.bandV1f); devFree(L.bandBf);
    L = Level();
}

// ====================================================================================================
// stand-alone domain builders
// ====================================================================================================
extern "C" int gmg_expand_dims(const int64_t baseRes[3], int64_t expRes[3], int64_t offset[3], int *mgLevels)
{
    if (!baseRes || !expRes || !offset || !mgLevels) return invalid("gmg_expand_dims: null argument");
    // HDK_GeometricMultigridOperators.h:1340-1360, reproduced literally (double log2 / ceil / pow / exp2)
    double minLog = std::min(std::log2(double(baseRes[0])), std::log2(double(baseRes[1])));
    minLog = std::min(minLog, std::log2(double(baseRes[2])));
    const int levels = int(std::ceil(minLog) - std::log2(2.0));
    const int pad = int(std::pow(2, levels - 1));
    for (int a = 0; a < 3; ++a)
    {
	const double logSize = std::ceil(std::log2(double(baseRes[a] + 2 * pad)));
	expRes[a] = int64_t(std::exp2(logSize));
	offset[a] = pad;
    }
    *mgLevels = levels;
    return GMG_OK;
}

extern "C" int gmg_expand_labels(gmg_ctx *ctx, const int32_t *base, const int64_t baseRes[3], int32_t *out, const int64_t expRes[3],
				 const int64_t offset[3])
{
    if (!ctx || !base || !out) return invalid("gmg_expand_labels: null argument");
    TRACE("enter");
    GMG_CUDA(enterCtx(ctx));
    TRACE("ctx set");
    const int64_t n = baseRes[0] * baseRes[1] * baseRes[2];
    int32_t *dIn = nullptr, *dOut = nullptr;
    GMG_CUDA(devMalloc(&dIn, sizeof(int32_t) * n));
    GMG_CUDA(devMalloc(&dOut, sizeof(int32_t) * n));
    TRACE("allocated");
    GMG_CUDA(cudaMemcpyAsync(dIn, base, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_expand_labels<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(dOut, dIn, n);
    }
    TRACE("kernel launched");
    // EXTERIOR everywhere, then the mapped base box at +offset
    std::fill(out, out + expRes[0] * expRes[1] * expRes[2], int32_t(L_EXTERIOR));
    TRACE("host filled");
    cudaMemcpy3DParms p = {};
    p.dstPtr = make_cudaPitchedPtr(out, size_t(expRes[0]) * 4, size_t(expRes[0]), size_t(expRes[1]));
    p.dstPos = make_cudaPos(size_t(offset[0]) * 4, size_t(offset[1]), size_t(offset[2]));
    p.srcPtr = make_cudaPitchedPtr(dOut, size_t(baseRes[0]) * 4, size_t(baseRes[0]), size_t(baseRes[1]));
    p.extent = make_cudaExtent(size_t(baseRes[0]) * 4, size_t(baseRes[1]), size_t(baseRes[2]));
    p.kind = cudaMemcpyDeviceToHost;
    GMG_CUDA(cudaMemcpy3DAsync(&p, ctx->stream));
    TRACE("3d copy issued");
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    TRACE("synced");
    GMG_CUDA(devFree(dIn));
    GMG_CUDA(devFree(dOut));
    return GMG_OK;
}

extern "C" int gmg_expand_weights(gmg_ctx *ctx, const double *baseW, const int64_t baseRes[3], double *out, const int64_t expRes[3],
				  const int64_t offset[3], int axis)
{
    if (!ctx || !baseW || !out || axis < 0 || axis > 2) return invalid("gmg_expand_weights: bad argument");
    GMG_CUDA(enterCtx(ctx));
    int64_t bfr[3] = {baseRes[0], baseRes[1], baseRes[2]}, efr[3] = {expRes[0], expRes[1], expRes[2]};
    ++bfr[axis];
    ++efr[axis];
    const int64_t n = bfr[0] * bfr[1] * bfr[2];
    double *dIn = nullptr, *dOut = nullptr;
    GMG_CUDA(devMalloc(&dIn, sizeof(double) * n));
    GMG_CUDA(devMalloc(&dOut, sizeof(double) * n));
    GMG_CUDA(cudaMemcpyAsync(dIn, baseW, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_expand_weights<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(dOut, dIn, n);
    }
    std::memset(out, 0, sizeof(double) * size_t(efr[0]) * efr[1] * efr[2]);
    cudaMemcpy3DParms p = {};
    p.dstPtr = make_cudaPitchedPtr(out, size_t(efr[0]) * 8, size_t(efr[0]), size_t(efr[1]));
    p.dstPos = make_cudaPos(size_t(offset[0]) * 8, size_t(offset[1]), size_t(offset[2]));
    p.srcPtr = make_cudaPitchedPtr(dOut, size_t(bfr[0]) * 8, size_t(bfr[0]), size_t(bfr[1]));
    p.extent = make_cudaExtent(size_t(bfr[0]) * 8, size_t(bfr[1]), size_t(bfr[2]));
    p.kind = cudaMemcpyDeviceToHost;
    GMG_CUDA(cudaMemcpy3DAsync(&p, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(dIn));
    GMG_CUDA(devFree(dOut));
    return GMG_OK;
}

template <typename LabelT>
static int boundsFromHintOrScan(const LabelT *labels, const int64_t res[3], const int64_t *hLo, const int64_t *hHi, int64_t lo[3], int64_t hi[3])
{
    const bool hint = hLo && hHi && (hHi[0] > hLo[0]) && (hHi[1] > hLo[1]) && (hHi[2] > hLo[2]);
    if (hint)
    {
	for (int a = 0; a < 3; ++a) { lo[a] = hLo[a]; hi[a] = hHi[a]; }
	return GMG_OK;
    }
    if (!scanBounds(labels, res, lo, hi))
    {
	setError("no non-EXTERIOR cell in the label grid");
	return GMG_ERR_NO_ACTIVE;
    }
    return GMG_OK;
}

// dW[a] are ALLOCATION bases covering the stored box plus one plane below and above (a slab's edge cells read the
// forward z-face weight of the next plane); kernels take dW[a] + g.plane.
static int uploadWeights(gmg_ctx *ctx, double *dW[3], const double *w0, const double *w1, const double *w2, const int64_t res[3], const Geom &g,
			 const int64_t *bounds = nullptr)
{
    const double *w[3] = {w0, w1, w2};
    Geom ge = g;
    ge.org[2] = g.org[2] - 1;
    ge.n[2] = g.n[2] + 2;
    ge.total = ge.plane * ge.n[2];
    for (int a = 0; a < 3; ++a)
    {
	int64_t fr[3] = {res[0], res[1], res[2]};
	++fr[a];
	GMG_CUDA(devMalloc(&dW[a], sizeof(double) * ge.total));
	int64_t fb[6];
	if (bounds)
	{
	    for (int k = 0; k < 6; ++k) fb[k] = bounds[k];
	    ++fb[3 + a];  // the face behind the last non-EXTERIOR cell
	}
	GMG_TRY(uploadValues(ctx, dW[a], w[a], fr, ge, nullptr, bounds ? fb : nullptr));
    }
    return GMG_OK;
}

extern "C" int gmg_set_boundary_labels(gmg_ctx *ctx, int32_t *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2,
				       const int64_t boxLo[3], const int64_t boxHi[3])
{
    if (!ctx || !labels || !w0 || !w1 || !w2) return invalid("gmg_set_boundary_labels: null argument");
    GMG_CUDA(enterCtx(ctx));
    int64_t lo[3], hi[3];
    int st = boundsFromHintOrScan(labels, res, boxLo, boxHi, lo, hi);
    if (st == GMG_ERR_NO_ACTIVE) return GMG_OK; // nothing to promote
    GMG_TRY(st);
    Geom g;
    makeGeom(g, res, lo, hi);
    uint8_t *dIn = nullptr, *dOut = nullptr;
    double *dW[3] = {nullptr, nullptr, nullptr};
    GMG_CUDA(devMalloc(&dIn, g.total));
    GMG_CUDA(devMalloc(&dOut, g.total));
    // the host arrays are only read inside the non-EXTERIOR box (a caller may hand over lazily allocated expanded grids)
    const int64_t bounds[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
    GMG_TRY(uploadLabels(ctx, dIn, labels, res, g, bounds));
    GMG_TRY(uploadWeights(ctx, dW, w0, w1, w2, res, g, bounds));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_set_boundary<<<unsigned(divUp(g.total, BLOCK)), BLOCK, 0, ctx->stream>>>(dOut, dIn, dW[0] + g.plane, dW[1] + g.plane, dW[2] + g.plane, boxArgs(g));
    }
    GMG_TRY(downloadLabels(ctx, labels, dOut, res, g, false, bounds));
    for (int a = 0; a < 3; ++a) GMG_CUDA(devFree(dW[a]));
    GMG_CUDA(devFree(dIn));
    GMG_CUDA(devFree(dOut));
    return GMG_OK;
}

// fine (level-like) geometry + its coarse geometry
static void coarseGeomOf(Geom &cg, int shift[3], const Geom &fg, const int64_t lo[3], const int64_t hi[3], int64_t clo[3], int64_t chi[3])
{
    int64_t cres[3];
    for (int a = 0; a < 3; ++a)
    {
	cres[a] = fg.res[a] / 2;
	clo[a] = floorDiv2(lo[a]);
	chi[a] = ceilDiv2(hi[a]);
    }
    makeGeom(cg, cres, clo, chi);
    for (int a = 0; a < 3; ++a) shift[a] = fg.org[a] / 2 - cg.org[a];
}

static int coarsenOnDevice(gmg_ctx *ctx, uint8_t *coarse, const Geom &cg, const uint8_t *fine, const Geom &fg, const int shift[3])
{
    uint8_t *tmp = nullptr;
    GMG_CUDA(devMalloc(&tmp, cg.total));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_coarsen1<<<unsigned(divUp(cg.total, BLOCK)), BLOCK, 0, ctx->stream>>>(tmp, fine, boxArgs(cg), boxArgs(fg), shift[0], shift[1], shift[2]);
    }
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_coarsen2<<<unsigned(divUp(cg.total, BLOCK)), BLOCK, 0, ctx->stream>>>(coarse, tmp, boxArgs(cg));
    }
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(devFree(tmp));
    return GMG_OK;
}

extern "C" int gmg_coarsen_labels(gmg_ctx *ctx, const int32_t *fine, const int64_t fineRes[3], int32_t *coarse)
{
    if (!ctx || !fine || !coarse) return invalid("gmg_coarsen_labels: null argument");
    for (int a = 0; a < 3; ++a)
	if (fineRes[a] % 2) return invalid("gmg_coarsen_labels: odd resolution");
    GMG_CUDA(enterCtx(ctx));
    int64_t lo[3], hi[3], clo[3], chi[3];
    const int64_t cres[3] = {fineRes[0] / 2, fineRes[1] / 2, fineRes[2] / 2};
    if (!scanBounds(fine, fineRes, lo, hi))
    {
	std::fill(coarse, coarse + cres[0] * cres[1] * cres[2], int32_t(L_EXTERIOR));
	return GMG_OK;
    }
    Geom fg, cg;
    int shift[3];
    makeGeom(fg, fineRes, lo, hi);
    coarseGeomOf(cg, shift, fg, lo, hi, clo, chi);
    uint8_t *dF = nullptr, *dC = nullptr;
    GMG_CUDA(devMalloc(&dF, fg.total));
    GMG_CUDA(devMalloc(&dC, cg.total));
    GMG_TRY(uploadLabels(ctx, dF, fine, fineRes, fg));
    GMG_TRY(coarsenOnDevice(ctx, dC, cg, dF, fg, shift));
    GMG_TRY(downloadLabels(ctx, coarse, dC, cres, cg, true));
    GMG_CUDA(devFree(dF));
    GMG_CUDA(devFree(dC));
    return GMG_OK;
}

extern "C" int gmg_boundary_cells(gmg_ctx *ctx, const int32_t *labels, const int64_t res[3], int width, int64_t *xyz, int64_t *count)
{
    if (!ctx || !labels || !count) return invalid("gmg_boundary_cells: null argument");
    GMG_CUDA(enterCtx(ctx));
    int64_t lo[3], hi[3];
    if (!scanBounds(labels, res, lo, hi)) { *count = 0; return GMG_OK; }
    Level L;
    makeGeom(L.g, res, lo, hi);
    GMG_CUDA(devMalloc(&L.labels, L.g.total));
    GMG_TRY(uploadLabels(ctx, L.labels, labels, res, L.g));
    L.gg = L.g;
    L.ownHi = L.g.n[2];
    GMG_TRY(buildBand(ctx, L, width));
    int st = exportBand(ctx, L, xyz, count);
    freeLevel(L);
    return st;
}

// ====================================================================================================
// solver construction (MG.cpp:135-418)
// ====================================================================================================
extern "C" void gmg_solver_default_options(gmg_solver_options *opt)
{
    if (!opt) return;
    std::memset(opt, 0, sizeof(*opt));
    opt->boundary_width = 3;
    opt->boundary_iterations = 3;
    opt->coarse_matrix_scale = 1.0;
}

// coarsest level: number the active cells like the reference (tile order, x fastest in tile; MG.cpp:296-323), assemble
// (MG.cpp:359-382), factor exactly (dense Cholesky) and keep the explicit inverse for a one-launch device solve.
static int buildCoarseSolve(gmg_solver *s)
{
    gmg_ctx *ctx = s->ctx;
    Level &L = s->lv[s->levels - 1];
    const Geom &g = L.g;
    std::vector<uint8_t> lab(g.total);
    GMG_CUDA(cudaMemcpy(lab.data(), L.labels, g.total, cudaMemcpyDeviceToHost));
    auto labelAtExp = [&](int64_t ex, int64_t ey, int64_t ez) -> int {
	const int64_t x = ex - g.org[0], y = ey - g.org[1], z = ez - g.org[2];
	if (x < 0 || y < 0 || z < 0 || x >= g.n[0] || y >= g.n[1] || z >= g.n[2]) return L_EXTERIOR;
	return lab[z * g.plane + y * g.pitch + x];
    };
    auto isActive = [](int l) { return l == L_INTERIOR || l == L_BOUNDARY; };
    // tile-ordered numbering over the stored box (tiles of the expanded grid)
    std::vector<int32_t> index(g.total, -1);
    std::vector<int32_t> cellIdx;
    const int64_t t0[3] = {std::max<int64_t>(0, g.org[0]) >> 4, std::max<int64_t>(0, g.org[1]) >> 4, std::max<int64_t>(0, g.org[2]) >> 4};
    const int64_t t1[3] = {(g.org[0] + g.n[0] + 15) >> 4, (g.org[1] + g.n[1] + 15) >> 4, (g.org[2] + g.n[2] + 15) >> 4};
    for (int64_t tz = t0[2]; tz < t1[2]; ++tz)
	for (int64_t ty = t0[1]; ty < t1[1]; ++ty)
	    for (int64_t tx = t0[0]; tx < t1[0]; ++tx)
		for (int64_t ez = tz * 16; ez < tz * 16 + 16; ++ez)
		    for (int64_t ey = ty * 16; ey < ty * 16 + 16; ++ey)
			for (int64_t ex = tx * 16; ex < tx * 16 + 16; ++ex)
			    if (isActive(labelAtExp(ex, ey, ez)))
			    {
				const int64_t si = (ez - g.org[2]) * g.plane + (ey - g.org[1]) * g.pitch + (ex - g.org[0]);
				index[si] = int32_t(cellIdx.size());
				cellIdx.push_back(int32_t(si));
			    }
    const int n = int(cellIdx.size());
    s->nCoarse = n;
    if (n == 0) { setError("coarsest level has no active cell"); return GMG_ERR_NO_ACTIVE; }
    if (n > MAX_COARSE)
    {
	char buf[256];
	snprintf(buf, sizeof(buf), "coarsest level has %d unknowns; the dense direct solve supports at most %d", n, MAX_COARSE);
	setError(buf);
	return GMG_ERR_COARSE_SIZE;
    }
    const double scale = s->opt.coarse_matrix_scale > 0 ? s->opt.coarse_matrix_scale : 1.0;
    std::vector<double> A(size_t(n) * n, 0.0);
    const int64_t stride[6] = {-1, 1, -int64_t(g.pitch), int64_t(g.pitch), -g.plane, g.plane};
    for (int r = 0; r < n; ++r)
    {
	const int64_t si = cellIdx[r];
	double diag = 0;
	for (int k = 0; k < 6; ++k)
	{
	    const int nl = lab[si + stride[k]];
	    if (isActive(nl)) { A[size_t(r) * n + index[si + stride[k]]] += -1.0 * scale; diag += 1; }
	    else if (nl == L_DIRICHLET) diag += 1;
	}
	A[size_t(r) * n + r] += diag * scale;
    }
    // Cholesky A = L L^T (in place, lower), then inverse via forward/back substitution on the identity
    std::vector<double> Lm(A);
    for (int j = 0; j < n; ++j)
    {
	double d = Lm[size_t(j) * n + j];
	for (int k = 0; k < j; ++k) d -= Lm[size_t(j) * n + k] * Lm[size_t(j) * n + k];
	if (!(d > 0))
	{
	    setError("coarse matrix is not positive definite (pure-Neumann domain without a DIRICHLET cell?)");
	    return GMG_ERR_NOT_SPD;
	}
	d = std::sqrt(d);
	Lm[size_t(j) * n + j] = d;
	for (int i = j + 1; i < n; ++i)
	{
	    double v = Lm[size_t(i) * n + j];
	    const double *ri = &Lm[size_t(i) * n], *rj = &Lm[size_t(j) * n];
	    for (int k = 0; k < j; ++k) v -= ri[k] * rj[k];
	    Lm[size_t(i) * n + j] = v / d;
	}
    }
    // Y = L^-1 (lower triangular), inv = Y^T Y
    std::vector<double> Y(size_t(n) * n, 0.0);
    for (int c = 0; c < n; ++c)
    {
	for (int i = c; i < n; ++i)
	{
	    double v = (i == c) ? 1.0 : 0.0;
	    const double *ri = &Lm[size_t(i) * n];
	    for (int k = c; k < i; ++k) v -= ri[k] * Y[size_t(k) * n + c];
	    Y[size_t(i) * n + c] = v / ri[i];
	}
    }
    std::vector<double> inv(size_t(n) * n, 0.0);
    for (int i = 0; i < n; ++i)
	for (int j = 0; j <= i; ++j)
	{
	    double v = 0;
	    for (int k = i; k < n; ++k) v += Y[size_t(k) * n + i] * Y[size_t(k) * n + j];
	    inv[size_t(i) * n + j] = v;
	    inv[size_t(j) * n + i] = v;
	}
    GMG_CUDA(devMalloc(&s->coarseIdx, sizeof(int32_t) * n));
    GMG_CUDA(devMalloc(&s->coarseInv, sizeof(double) * size_t(n) * n));
    GMG_CUDA(cudaMemcpy(s->coarseIdx, cellIdx.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    GMG_CUDA(cudaMemcpy(s->coarseInv, inv.data(), sizeof(double) * size_t(n) * n, cudaMemcpyHostToDevice));
    return GMG_OK;
}

// ====================================================================================================
// z-slab sharding
// ====================================================================================================
// Levels [0, S) are cut into z-slabs, one per rank, with cut planes that nest from level to level (a coarse plane and
// its two child planes always belong to the same rank); levels [S, L) are replicated.  Every rank stores its owned
// planes plus a deep halo and recomputes the halo redundantly, so a level needs two plane exchanges per V-cycle (rhs on
// the way down, solution on the way up) instead of one per sweep; each cell is computed with the same arithmetic as on
// one GPU, so sharded results are bitwise equal to unsharded ones apart from the order of the dot-product partial sums.
extern "C" int gmg_shard_plan(const int64_t *levelPlanes, const int64_t *levelShiftZ, const int64_t *levelCells, int levels, int world,
			      int maxShardLevels, int64_t minCells, int *shardLevels, int64_t *cuts)
{
    // levelPlanes[l]: z extent of level l's global storage box; levelShiftZ[l]: coarse plane = (fine plane >> 1) + shift;
    // cuts: [levels][world + 1] global storage planes (filled for the sharded levels)
    if (!levelPlanes || !levelShiftZ || !levelCells || !shardLevels || !cuts || levels < 1 || world < 1) return invalid("gmg_shard_plan: bad argument");
    int S = 0;
    if (world > 1)
	for (int l = 0; l < std::min(levels - 1, maxShardLevels); ++l)
	{
	    if (levelCells[l] < minCells) break;
	    S = l + 1;
	}
    for (; S > 0; --S)
    {
	bool ok = true;
	const int64_t nS = levelPlanes[S - 1];
	int64_t *c = cuts + int64_t(S - 1) * (world + 1);
	for (int k = 0; k <= world; ++k) c[k] = 2 * int64_t(std::llround(double(k) * double(nS) / (2.0 * world)));
	c[0] = 0;
	c[world] = nS;
	for (int l = S - 2; l >= 0; --l)
	{
	    const int64_t *cc = cuts + int64_t(l + 1) * (world + 1);
	    int64_t *cf = cuts + int64_t(l) * (world + 1);
	    for (int k = 1; k < world; ++k) cf[k] = std::max<int64_t>(0, std::min<int64_t>(levelPlanes[l], 2 * (cc[k] - levelShiftZ[l])));
	    cf[0] = 0;
	    cf[world] = levelPlanes[l];
	}
	for (int l = 0; l < S && ok; ++l)
	{
	    const int64_t need = (l == 0) ? HALO_STORE0 : HALO_STORE;
	    const int64_t *cl = cuts + int64_t(l) * (world + 1);
	    for (int k = 0; k < world; ++k)
		if (cl[k + 1] - cl[k] < need || (cl[k] & 1)) ok = false;
	}
	if (ok) break;
    }
    *shardLevels = S;
    return GMG_OK;
}

static int planShards(gmg_solver *s)
{
    gmg_ctx *ctx = s->ctx;
    const int nl = s->levels, world = ctx->world, rank = ctx->rank;
    for (auto &L : s->lv)
    {
	L.gg = L.g;
	L.labelsAlloc = L.labels;
	L.sharded = false;
	L.zOff = 0;
	L.ownLo = 0;
	L.ownHi = L.g.n[2];
    }
    s->shardLevels = 0;
    if (world == 1) return GMG_OK;
    if (s->opt.boundary_iterations > 3) return invalid("sharded contexts support at most 3 boundary smoother iterations (halo depth)");
    int maxS = 3;
    int64_t minCells = 1500000;
    if (const char *e = getenv("GMG_SHARD_LEVELS")) maxS = atoi(e);
    if (const char *e = getenv("GMG_SHARD_MIN_CELLS")) minCells = atoll(e);
    std::vector<int64_t> planes(nl), shiftZ(nl), cells(nl), cuts(size_t(nl) * (world + 1), 0);
    for (int l = 0; l < nl; ++l) { planes[l] = s->lv[l].g.n[2]; shiftZ[l] = s->lv[l].shift[2]; cells[l] = s->lv[l].g.total; }
    int S = 0;
    GMG_TRY(gmg_shard_plan(planes.data(), shiftZ.data(), cells.data(), nl, world, maxS, minCells, &S, cuts.data()));
    s->shardLevels = S;
    if (S == 0) return GMG_OK;
    for (int l = 0; l < S; ++l)
    {
	Level &L = s->lv[l];
	const int64_t *c = cuts.data() + size_t(l) * (world + 1);
	const int H = (l == 0) ? HALO_STORE0 : HALO_STORE;
	const int zs = int(std::max<int64_t>(0, c[rank] - H)), ze = int(std::min<int64_t>(L.gg.n[2], c[rank + 1] + H));
	L.sharded = true;
	L.zOff = zs;
	L.ownLo = int(c[rank]) - zs;
	L.ownHi = int(c[rank + 1]) - zs;
	L.g.org[2] = L.gg.org[2] + zs;
	L.g.n[2] = ze - zs;
	L.g.total = L.g.plane * L.g.n[2];
	L.g.zBlocks = int(divUp(L.g.n[2], CHUNK_Z));
	L.labels = L.labelsAlloc + int64_t(zs) * L.g.plane;
	int64_t nI = 0, nB = 0;
	GMG_TRY(countLabels(ctx, L.labels, L.g.total, &nI, &nB));
	L.nInterior = nI;
	L.nActive = nI + nB;
    }
    // rows of a plane that hold an active cell anywhere in the level (global labels: identical on every rank): the halo
    // exchange moves only those (haloExchange / k_halo_p2p)
    for (int l = 0; l < S; ++l)
    {
	Level &L = s->lv[l];
	const int nzG = L.gg.n[2];
	int4 *d = nullptr;
	GMG_CUDA(devMalloc(&d, sizeof(int4) * nzG));
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_plane_extents<<<nzG, BLOCK, 0, ctx->stream>>>(d, L.labelsAlloc, L.gg.n[0], L.gg.n[1], L.gg.pitch, L.gg.plane);
	}
	std::vector<int4> e(nzG);
	GMG_CUDA(cudaMemcpyAsync(e.data(), d, sizeof(int4) * nzG, cudaMemcpyDeviceToHost, ctx->stream));
	GMG_CUDA(cudaStreamSynchronize(ctx->stream));
	GMG_CUDA(devFree(d));
	int y0 = L.gg.n[1], y1 = 0;
	for (const int4 &q : e)
	    if (q.y > q.x) { y0 = std::min(y0, q.z); y1 = std::max(y1, q.w); }
	if (y1 <= y0) { y0 = 0; y1 = L.gg.n[1]; }
	static const bool whole = [] { const char *e = getenv("GMG_HALO_ROWS"); return e && e[0] == '0'; }();
	L.rowLo = whole ? 0 : y0;
	L.rowHi = whole ? L.gg.n[1] : y1;
    }
    // fine -> coarse storage relation between the LOCAL boxes
    for (int l = 0; l + 1 < nl; ++l)
	for (int a = 0; a < 3; ++a) s->lv[l].shift[a] = s->lv[l].g.org[a] / 2 - s->lv[l + 1].g.org[a];
    // planes of the first replicated level each rank restricts into (they tile that level's box)
    {
	const int64_t *c = cuts.data() + size_t(S - 1) * (world + 1);
	const int64_t nC = s->lv[S].gg.n[2];
	s->gatherLo.assign(world, 0);
	s->gatherHi.assign(world, 0);
	for (int k = 0; k < world; ++k)
	{
	    s->gatherLo[k] = (k == 0) ? 0 : int(std::min<int64_t>(nC, std::max<int64_t>(0, (c[k] >> 1) + shiftZ[S - 1])));
	    s->gatherHi[k] = (k == world - 1) ? int(nC) : int(std::min<int64_t>(nC, std::max<int64_t>(0, (c[k + 1] >> 1) + shiftZ[S - 1])));
	}
    }
    return GMG_OK;
}

// ---- peer-memory communication (gmg_p2p.cuh) ------------------------------------------------------------------------
static void p2pRelease(gmg_ctx *ctx)
{
    P2pState *st = static_cast<P2pState *>(ctx->p2p);
    if (!st) return;
    for (int r = 0; r < st->world; ++r)
	if (r != st->rank && st->peer[r]) cudaIpcCloseMemHandle(st->peer[r]);
    cudaFree(st->arena);
    cudaFree(st->seq);
    cudaFree(st->tickets);
    cudaFree(st->error);
    delete st;
    ctx->p2p = nullptr;
}

// mailboxes for this solver's sharded levels; (re)allocates and re-maps the arenas when they are too small.  Collective.
static int p2pEnsure(gmg_solver *s)
{
    gmg_ctx *ctx = s->ctx;
    if (ctx->world == 1 || s->shardLevels == 0 || ctx->p2pDisabled) return GMG_OK;
    if (const char *e = getenv("GMG_P2P")) if (e[0] == '0') { ctx->p2pDisabled = true; return GMG_OK; }
    if (ctx->world > P2P_MAX_WORLD || s->shardLevels > P2P_MAX_LEVELS) { ctx->p2pDisabled = true; return GMG_OK; }
    const NcclApi *api = ncclApi(nullptr);
    NcclComm comm = static_cast<NcclComm>(ctx->nccl);
    // Box capacities only ever grow within a context: ONE layout per arena generation, shared by every live solver, so two
    // solvers of different sizes can never disagree about where a mailbox lives (a solver's planes just fill part of a box).
    size_t needHalo[P2P_MAX_LEVELS] = {0, 0, 0, 0};
    for (int l = 0; l < s->shardLevels; ++l) needHalo[l] = sizeof(double) * size_t(l == 0 ? HALO_STORE0 : HALO_STORE) * size_t(s->lv[l].g.plane);
    const size_t needGather = sizeof(double) * size_t(s->lv[s->shardLevels].g.total);
    P2pState *st = static_cast<P2pState *>(ctx->p2p);
    bool fits = st != nullptr;
    for (int l = 0; fits && l < P2P_MAX_LEVELS; ++l) fits = needHalo[l] <= st->haloCap[l];
    if (fits && needGather <= st->gatherCap)
    {
	s->p2pGeneration = st->generation;
	return GMG_OK;
    }
    size_t capHalo[P2P_MAX_LEVELS], capGather;
    auto grow = [](size_t need, size_t have) { return need <= have ? have : ((need + (need >> 2) + 255) & ~size_t(255)); };  // head-room: the next frame's boxes differ by a few planes
    for (int l = 0; l < P2P_MAX_LEVELS; ++l) capHalo[l] = grow(needHalo[l], st ? st->haloCap[l] : 0);
    capGather = grow(needGather, st ? st->gatherCap : 0);
    P2pLayout lay;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    lay.flags = take(sizeof(unsigned long long) * P2P_FLAG_COUNT);
    lay.scalars = take(sizeof(double) * 2 * P2P_MAX_WORLD);
    for (int l = 0; l < P2P_MAX_LEVELS; ++l)
	for (int d = 0; d < 2; ++d)
	    for (int k = 0; k < 2; ++k) lay.halo[l][d][k] = take(capHalo[l]);
    for (int k = 0; k < 2; ++k) lay.gather[k] = take(capGather);
    lay.bytes = off;
    // (re)build: everybody drains, frees, allocates, exchanges IPC handles
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_NCCL(api->AllReduce(ctx->scalars, ctx->scalars, 1, NCCL_FLOAT64, NCCL_SUM, comm, ctx->stream));  // barrier
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    // cached graphs of the solvers that are still alive point into the arenas about to be freed: drop them (they are
    // re-captured against the new arenas on their next use; the grown layout holds every older solver's planes too)
    for (gmg_solver *o : ctx->solvers) dropGraphs(o);
    p2pRelease(ctx);
    st = new P2pState;
    st->rank = ctx->rank;
    st->world = ctx->world;
    st->generation = ++ctx->p2pGenerations;
    s->p2pGeneration = st->generation;
    st->layout = lay;
    for (int l = 0; l < P2P_MAX_LEVELS; ++l) st->haloCap[l] = capHalo[l];
    st->gatherCap = capGather;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    std::vector<cudaIpcMemHandle_t> all(ctx->world);
    unsigned char *dh = nullptr;
    if (cudaMalloc(&st->arena, st->layout.bytes) != cudaSuccess) ok = 0;
    if (ok) ok = cudaMemset(st->arena, 0, st->layout.bytes) == cudaSuccess;
    if (ok) ok = cudaMalloc(&st->seq, sizeof(unsigned long long) * (P2P_MAX_LEVELS + 2)) == cudaSuccess && cudaMalloc(&st->tickets, 2 * sizeof(unsigned)) == cudaSuccess &&
		 cudaMalloc(&st->error, sizeof(int)) == cudaSuccess;
    if (ok)
    {
	cudaMemset(st->seq, 0, sizeof(unsigned long long) * (P2P_MAX_LEVELS + 2));
	cudaMemset(st->tickets, 0, 2 * sizeof(unsigned));
	cudaMemset(st->error, 0, sizeof(int));
	ok = cudaIpcGetMemHandle(&mine, st->arena) == cudaSuccess;
    }
    cudaGetLastError();
    if (!ok) std::memset(&mine, 0, sizeof(mine));
    GMG_CUDA(cudaMalloc(&dh, sizeof(cudaIpcMemHandle_t) * ctx->world));
    GMG_CUDA(cudaMemcpy(dh + sizeof(mine) * ctx->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    GMG_NCCL(api->GroupStart());
    for (int r = 0; r < ctx->world; ++r)
	GMG_NCCL(api->Broadcast(dh + sizeof(mine) * r, dh + sizeof(mine) * r, sizeof(mine), NCCL_UINT8, r, comm, ctx->stream));
    GMG_NCCL(api->GroupEnd());
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(cudaMemcpy(all.data(), dh, sizeof(mine) * ctx->world, cudaMemcpyDeviceToHost));
    cudaFree(dh);
    st->peer[ctx->rank] = st->arena;
    for (int r = 0; r < ctx->world && ok; ++r)
    {
	if (r == ctx->rank) continue;
	void *p = nullptr;
	if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
	st->peer[r] = static_cast<char *>(p);
    }
    // all or nothing: one rank that cannot map its peers sends everybody back to NCCL
    double flag = ok ? 0.0 : 1.0, *dflag = scalarPtr(s, offsetof(Scalars, tmp));
    GMG_CUDA(cudaMemcpy(dflag, &flag, sizeof(double), cudaMemcpyHostToDevice));
    GMG_NCCL(api->AllReduce(dflag, dflag, 1, NCCL_FLOAT64, NCCL_SUM, comm, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(cudaMemcpy(&flag, dflag, sizeof(double), cudaMemcpyDeviceToHost));
    ctx->p2p = st;
    if (flag != 0.0)
    {
	p2pRelease(ctx);
	ctx->p2pDisabled = true;
	if (getenv("GMG_TRACE")) fprintf(stderr, "[gmg] peer-memory mapping failed on some rank: NCCL for every exchange\n");
    }
    return GMG_OK;
}

static int p2pCheckError(gmg_ctx *ctx)
{
    P2pState *st = static_cast<P2pState *>(ctx->p2p);
    if (!st) return GMG_OK;
    int e = 0;
    GMG_CUDA(cudaMemcpy(&e, st->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (e)
    {
	setError("peer-memory exchange timed out waiting for a neighbour rank (ranks must make the same calls in the same order)");
	return GMG_ERR_COMM;
    }
    return GMG_OK;
}

// the context's arenas, for a solver whose planes were sized into them (capacities never shrink, so any generation will do)
static P2pState *p2pOf(gmg_solver *s)
{
    P2pState *st = static_cast<P2pState *>(s->ctx->p2p);
    return (st && s->p2pGeneration >= 0) ? st : nullptr;
}

struct ZRange { int lo, hi; };
// planes within `depth` of the owned slab (everything on an unsharded level)
static ZRange clipDepth(const Level &L, int depth)
{
    if (!L.sharded) return {0, L.g.n[2]};
    return {std::max(0, L.ownLo - depth), std::min(L.g.n[2], L.ownHi + depth)};
}

// refresh `depth` halo planes of grid p on a sharded level from the two z-neighbours (NCCL P2P over NVLink)
static int haloExchange(gmg_solver *s, int level, double *p, int depth)
{
    gmg_ctx *ctx = s->ctx;
    const Level &L = s->lv[level];
    if (!L.sharded || ctx->world == 1 || depth <= 0) return GMG_OK;
    const NcclApi *api = ncclApi(nullptr);
    NcclComm comm = static_cast<NcclComm>(ctx->nccl);
    const size_t cnt = size_t(depth) * size_t(L.g.plane);
    const int rank = ctx->rank, world = ctx->world;
    ctx->curLevel = level;
    GMG_LAUNCH(ctx, KC_HALO, double(cnt) * 8.0 * ((rank > 0) + (rank < world - 1)) * 2.0);
    if (P2pState *st = p2pOf(s))
    {
	const P2pLayout &lay = st->layout;
	HaloP2pArgs a;
	a.grid = p; a.plane = L.g.plane; a.ownLo = L.ownLo; a.ownHi = L.ownHi; a.depth = depth;
	a.segOff = int64_t(L.rowLo) * L.g.pitch;
	a.segLen = int64_t(L.rowHi - L.rowLo) * L.g.pitch;
	a.hasLower = rank > 0; a.hasUpper = rank < world - 1;
	a.slotStride = int64_t((lay.halo[level][0][1] - lay.halo[level][0][0]) / sizeof(double));
	unsigned long long *myFlags = reinterpret_cast<unsigned long long *>(st->arena + lay.flags);
	// box / flag index: dir 0 = data that came from the lower neighbour, dir 1 = from the upper neighbour
	a.fromLower = reinterpret_cast<const double *>(st->arena + lay.halo[level][0][0]);
	a.fromUpper = reinterpret_cast<const double *>(st->arena + lay.halo[level][1][0]);
	a.myFlagLower = myFlags + P2P_FLAG_HALO + (level * 2 + 0) * 2;
	a.myFlagUpper = myFlags + P2P_FLAG_HALO + (level * 2 + 1) * 2;
	a.toLower = a.toUpper = nullptr; a.flagOnLower = a.flagOnUpper = nullptr;
	if (a.hasLower)
	{
	    a.toLower = reinterpret_cast<double *>(st->peer[rank - 1] + lay.halo[level][1][0]);  // I am its upper neighbour
	    a.flagOnLower = reinterpret_cast<unsigned long long *>(st->peer[rank - 1] + lay.flags) + P2P_FLAG_HALO + (level * 2 + 1) * 2;
	}
	if (a.hasUpper)
	{
	    a.toUpper = reinterpret_cast<double *>(st->peer[rank + 1] + lay.halo[level][0][0]);  // I am its lower neighbour
	    a.flagOnUpper = reinterpret_cast<unsigned long long *>(st->peer[rank + 1] + lay.flags) + P2P_FLAG_HALO + (level * 2 + 0) * 2;
	}
	a.seq = st->seq + level; a.tickets = st->tickets; a.error = st->error;
	// small messages: fewer CTAs meet at the tickets and poll the flags (one CTA per 8 KB of a direction's payload, at least 8)
	const size_t payload = size_t(depth) * size_t(a.segLen) * sizeof(double);
	const unsigned ctas = unsigned(std::max<size_t>(8, std::min<size_t>(P2P_CTAS, payload / 8192)));
	k_halo_p2p<<<ctas, 256, 0, ctx->stream>>>(a);
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    GMG_NCCL(api->GroupStart());
    if (rank > 0)
    {
	GMG_NCCL(api->Send(p + int64_t(L.ownLo) * L.g.plane, cnt, NCCL_FLOAT64, rank - 1, comm, ctx->stream));
	GMG_NCCL(api->Recv(p + int64_t(L.ownLo - depth) * L.g.plane, cnt, NCCL_FLOAT64, rank - 1, comm, ctx->stream));
    }
    if (rank < world - 1)
    {
	GMG_NCCL(api->Send(p + int64_t(L.ownHi - depth) * L.g.plane, cnt, NCCL_FLOAT64, rank + 1, comm, ctx->stream));
	GMG_NCCL(api->Recv(p + int64_t(L.ownHi) * L.g.plane, cnt, NCCL_FLOAT64, rank + 1, comm, ctx->stream));
    }
    GMG_NCCL(api->GroupEnd());
    return GMG_OK;
}

// all-gather of the first replicated level's grid: every rank contributes the planes it restricted into
static int gatherReplicated(gmg_solver *s, double *p)
{
    gmg_ctx *ctx = s->ctx;
    if (ctx->world == 1 || s->shardLevels == 0) return GMG_OK;
    const Level &C = s->lv[s->shardLevels];
    const NcclApi *api = ncclApi(nullptr);
    NcclComm comm = static_cast<NcclComm>(ctx->nccl);
    ctx->curLevel = s->shardLevels;
    GMG_LAUNCH(ctx, KC_HALO, double(C.g.total) * 8.0);
    if (P2pState *st = p2pOf(s))
    {
	const P2pLayout &lay = st->layout;
	GatherP2pArgs a;
	a.grid = p; a.plane = C.g.plane; a.rank = ctx->rank; a.world = ctx->world;
	for (int k = 0; k < ctx->world; ++k)
	{
	    a.lo[k] = s->gatherLo[k]; a.hi[k] = s->gatherHi[k];
	    a.peerBox[k] = reinterpret_cast<double *>(st->peer[k] + lay.gather[0]);
	    a.flagOnPeer[k] = reinterpret_cast<unsigned long long *>(st->peer[k] + lay.flags) + P2P_FLAG_GATHER;
	}
	a.slotStride = int64_t((lay.gather[1] - lay.gather[0]) / sizeof(double));
	a.myBox = reinterpret_cast<const double *>(st->arena + lay.gather[0]);
	a.myFlags = reinterpret_cast<const unsigned long long *>(st->arena + lay.flags) + P2P_FLAG_GATHER;
	a.seq = st->seq + P2P_MAX_LEVELS; a.tickets = st->tickets; a.error = st->error;
	const unsigned ctas = unsigned(std::max<size_t>(8, std::min<size_t>(P2P_CTAS, size_t(C.g.total) * sizeof(double) / 65536)));
	k_gather_p2p<<<ctas, 256, 0, ctx->stream>>>(a);
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    GMG_NCCL(api->GroupStart());
    for (int k = 0; k < ctx->world; ++k)
    {
	const size_t cnt = size_t(s->gatherHi[k] - s->gatherLo[k]) * size_t(C.g.plane);
	if (cnt == 0) continue;
	double *q = p + int64_t(s->gatherLo[k]) * C.g.plane;
	GMG_NCCL(api->Broadcast(q, q, cnt, NCCL_FLOAT64, k, comm, ctx->stream));
    }
    GMG_NCCL(api->GroupEnd());
    return GMG_OK;
}

// sum of one device double over the ranks (CG dot products), in place
static int allreduceScalar(gmg_solver *s, double *dev, int op = NCCL_SUM)
{
    gmg_ctx *ctx = s->ctx;
    if (ctx->world == 1 || s->shardLevels == 0) return GMG_OK;
    const NcclApi *api = ncclApi(nullptr);
    ctx->curLevel = 0;
    GMG_LAUNCH(ctx, KC_HALO, 8.0);
    if (P2pState *st = p2pOf(s))
    {
	const P2pLayout &lay = st->layout;
	ScalarP2pArgs a;
	a.value = dev; a.rank = ctx->rank; a.world = ctx->world; a.isMax = (op == NCCL_MAX);
	for (int k = 0; k < ctx->world; ++k)
	{
	    a.peerSlots[k] = reinterpret_cast<double *>(st->peer[k] + lay.scalars);
	    a.flagOnPeer[k] = reinterpret_cast<unsigned long long *>(st->peer[k] + lay.flags) + P2P_FLAG_SCALAR;
	}
	a.mySlots = reinterpret_cast<const double *>(st->arena + lay.scalars);
	a.myFlags = reinterpret_cast<const unsigned long long *>(st->arena + lay.flags) + P2P_FLAG_SCALAR;
	a.seq = st->seq + P2P_MAX_LEVELS + 1; a.error = st->error;
	k_scalar_p2p<<<1, 32, 0, ctx->stream>>>(a);
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    GMG_NCCL(api->AllReduce(dev, dev, 1, NCCL_FLOAT64, op, static_cast<NcclComm>(ctx->nccl), ctx->stream));
    return GMG_OK;
}

static int buildFusedCycle(gmg_solver *s);
static int buildClusterSmooth(gmg_solver *s);
static int buildBandTiles(gmg_solver *s);

extern "C" int gmg_solver_destroy(gmg_solver *s)
{
    if (!s) return GMG_OK;
    enterCtx(s->ctx);
    cudaStreamSynchronize(s->ctx->stream);
    dropGraphs(s);
    if (s->ctx->windowOwner == s)
    {
	cudaStreamAttrValue attr = {};
	cudaStreamSetAttribute(s->ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
	s->ctx->windowOwner = nullptr;
    }
    {
	auto &v = s->ctx->solvers;
	v.erase(std::remove(v.begin(), v.end(), s), v.end());
    }
    devFree(s->coarseIdx); devFree(s->coarseInv); devFree(s->compactBlob); devFree(s->pcgLoop);
    if (s->pcgLoopHost) cudaFreeHost(s->pcgLoopHost);
    delete static_cast<CompactArgs *>(s->compactArgs);
    delete static_cast<ClusterArgs *>(s->clusterArgs);
    devFree(s->clusterSlab);
    if (!s->lv.empty())
    {
	const Geom &g0 = s->lv[0].g;
	freeGrid(s->pcgR, g0); freeGrid(s->pcgP, g0); freeGrid(s->pcgZ, g0); freeGrid(s->pcgT, g0); freeGrid(s->pcgX, g0); freeGrid(s->pcgB, g0); freeGrid(s->diagInv, g0);
    }
    for (auto &L : s->lv) freeLevel(L);
    delete s;
    return GMG_OK;
}

template <typename LabelT>
static int solverCreate(gmg_ctx *ctx, const LabelT *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2,
			int mgLevels, const gmg_solver_options *optIn, gmg_solver **out)
{
    if (!ctx || !labels || !res || !out) return invalid("gmg_solver_create: null argument");
    if ((w0 || w1 || w2) && !(w0 && w1 && w2)) return invalid("gmg_solver_create: pass all three weight grids or none");
    if (mgLevels < 1) return invalid("gmg_solver_create: mgLevels must be >= 1");
    for (int a = 0; a < 3; ++a)
    {
	if (res[a] % 2) return invalid("gmg_solver_create: resolution must be even (MG.cpp:155-157)");
	if (int(std::log2(double(res[a]))) + 1 < mgLevels) return invalid("gmg_solver_create: too many levels for this resolution (MG.cpp:159-161)");
	if ((res[a] >> (mgLevels - 1)) << (mgLevels - 1) != res[a]) return invalid("gmg_solver_create: resolution not divisible by 2^(levels-1)");
    }
    GMG_CUDA(enterCtx(ctx));
    const double tStart = nowMs();
    gmg_solver *s = new gmg_solver;
    s->ctx = ctx;
    ctx->solvers.push_back(s);
    if (optIn) s->opt = *optIn;
    else gmg_solver_default_options(&s->opt);
    if (const char *e = getenv("GMG_NO_GRAPHS")) s->useGraphs = !(e[0] == '1');
    if (const char *e = getenv("GMG_PRINT_STATS")) s->opt.print_stats = (e[0] == '1');
    if (const char *e = getenv("GMG_ZERO_AWARE")) s->zeroAware = !(e[0] == '0');
    if (const char *e = getenv("GMG_MIXED")) s->opt.mixed_precision = (e[0] == '1');
    s->tmaMask = tmaMode();
    if (const char *e = getenv("GMG_BAND_GROUPS")) s->bandGroups = (e[0] == '1');
    if (const char *e = getenv("GMG_BAND_RESIDENT")) s->bandResident = (e[0] == '1');
    if (const char *e = getenv("GMG_BAND_TILES")) s->bandTiles = (e[0] == '1');
    if (const char *e = getenv("GMG_BAND_GROUP_MAX")) s->bandGroupMax = atoll(e);
    if (const char *e = getenv("GMG_BAND_PER_THREAD")) s->bandPerThread = atoi(e);
    if (const char *e = getenv("GMG_STENCIL_CAP")) s->stencilCap = atoi(e);
    if (const char *e = getenv("GMG_STENCIL_BATCH")) s->stencilBatch = atoi(e);
    if (const char *e = getenv("GMG_STENCIL_LOOP")) s->stencilLoop = atoi(e);
    if (s->opt.boundary_width < 1) s->opt.boundary_width = 3;
    if (s->opt.boundary_iterations < 0) s->opt.boundary_iterations = 3;
    if (s->opt.use_gauss_seidel && ctx->world > 1)
    {
	ctx->solvers.pop_back();
	delete s;
	return invalid("gmg_solver_create: the tiled Gauss-Seidel smoother is not available on a sharded context (its in-tile dependencies reach 16 "
		       "cells, beyond the deep halo); use the damped-Jacobi smoother there");
    }
    auto fail = [&](int st) { gmg_solver_destroy(s); return st; };
    double tPhase = tStart;
    auto lap = [&](const char *name) {
	if (!s->opt.print_stats) return;
	cudaStreamSynchronize(ctx->stream);
	const double t = nowMs();
	printf("      %-28s %8.3f ms\n", name, t - tPhase);
	tPhase = t;
    };

    int64_t lo[3], hi[3];
    int st = boundsFromHintOrScan(labels, res, s->opt.box_lo, s->opt.box_hi, lo, hi);
    if (st != GMG_OK) return fail(st);

    s->lv.resize(mgLevels);
    s->levels = mgLevels;
    double *dW[3] = {nullptr, nullptr, nullptr};
    CoefJob coefJob;  // joins in its destructor on every early return
    {
	Level &L0 = s->lv[0];
	makeGeom(L0.g, res, lo, hi);
	if (L0.g.total >= (int64_t(1) << 31)) return fail(invalid("gmg_solver_create: cropped box exceeds 2^31 cells"));
	if ((st = (devMalloc(&L0.labels, L0.g.total) == cudaSuccess ? GMG_OK : GMG_ERR_CUDA)) != GMG_OK) return fail(st);
	for (int a = 0; a < 3; ++a) { s->hostBounds[a] = lo[a]; s->hostBounds[3 + a] = hi[a]; }
	if ((st = uploadLabels(ctx, L0.labels, labels, res, L0.g, s->hostBounds)) != GMG_OK) return fail(st);
    }
    lap("upload labels");
    // Single GPU with face weights: the level-0 band and the host gather of its face weights start right away, so the
    // gather (a few ms of cache misses in the caller's GB-sized weight grids) runs beside the whole rest of the constructor
    static const bool fullWeights = [] { const char *e = getenv("GMG_FULL_WEIGHTS"); return e && e[0] == '1'; }();
    bool earlyBand0 = false;
    if (ctx->world == 1 && w0 && !fullWeights)
    {
	Level &L0 = s->lv[0];
	L0.gg = L0.g;
	L0.labelsAlloc = L0.labels;
	L0.ownLo = 0;
	L0.ownHi = L0.g.n[2];
	if ((st = buildBand(ctx, L0, s->opt.boundary_width)) != GMG_OK) return fail(st);
	if ((st = buildCoefsSparse(ctx, L0, w0, w1, w2, res, s->hostBounds, &coefJob)) != GMG_OK) return fail(st);
	earlyBand0 = true;
	lap("level 0 band, face-weight gather started");
    }
    // coarse labels (MG.cpp:238-253) over the GLOBAL box of every level (one byte per cell, replicated on every rank),
    // with the reference's level cap: a level without active cells drops it AND the one before
    int64_t clo[3] = {lo[0], lo[1], lo[2]}, chi[3] = {hi[0], hi[1], hi[2]};
    for (int level = 0; level < s->levels; ++level)
    {
	Level &L = s->lv[level];
	if (level > 0)
	{
	    Level &F = s->lv[level - 1];
	    int64_t nlo[3], nhi[3];
	    coarseGeomOf(L.g, F.shift, F.g, clo, chi, nlo, nhi);
	    for (int a = 0; a < 3; ++a) { clo[a] = nlo[a]; chi[a] = nhi[a]; }
	    if (devMalloc(&L.labels, L.g.total) != cudaSuccess) return fail(GMG_ERR_CUDA);
	    if ((st = coarsenOnDevice(ctx, L.labels, L.g, F.labels, F.g, F.shift)) != GMG_OK) return fail(st);
	}
	int64_t nI = 0, nB = 0;
	if ((st = countLabels(ctx, L.labels, L.g.total, &nI, &nB)) != GMG_OK) return fail(st);
	L.nInterior = nI;
	L.nActive = nI + nB;
	L.nActiveGlobal = nI + nB;
	if (L.nActive == 0)
	{
	    if (level == 0) { setError("no INTERIOR/BOUNDARY cell at level 0"); return fail(GMG_ERR_NO_ACTIVE); }
	    const int newLevels = level - 1; // MG.cpp:245
	    for (int l = std::max(newLevels, 0); l < int(s->lv.size()); ++l) freeLevel(s->lv[l]);
	    s->levels = newLevels;
	    s->lv.resize(std::max(newLevels, 0));
	    break;
	}
    }
    if (s->levels < 1)
    {
	setError("level cap (MG.cpp:243-248) left no level: level 1 has no active cell");
	return fail(GMG_ERR_NO_ACTIVE);
    }
    lap("coarse labels");
    // z-slab views of the fine levels (world > 1); afterwards L.g / L.labels are the rank's local box
    if ((st = planShards(s)) != GMG_OK) return fail(st);
    // no weight grids = the reference's boundaryWeights == nullptr form (weight 1 to active/DIRICHLET neighbours, Ops.h:237-248)
    // the weight grids are NOT uploaded: only BOUNDARY cells of level 0 look at face weights (buildCoefsSparse); GMG_FULL_WEIGHTS=1
    // restores the upload of the three grids
    if (w0 && fullWeights && (st = uploadWeights(ctx, dW, w0, w1, w2, res, s->lv[0].g, s->hostBounds)) != GMG_OK) return fail(st);
    lap("shard plan");
    // bands (MG.cpp:279-281), coefficient records, chunk lists, grids
    int maxGrid = 0;
    for (int level = 0; level < s->levels; ++level)
    {
	Level &L = s->lv[level];
	const bool early = level == 0 && earlyBand0;
	if (!early && (st = buildBand(ctx, L, s->opt.boundary_width)) != GMG_OK) return fail(st);
	const int64_t wOff = L.g.plane;
	const bool fw = level == 0 && dW[0];
	if (level == 0) lap("level 0 band");
	if (early) st = GMG_OK;
	else if (level == 0 && w0 && !fullWeights) st = buildCoefsSparse(ctx, L, w0, w1, w2, res, s->hostBounds, &coefJob);
	else st = buildCoefs(ctx, L, fw ? dW[0] + wOff : nullptr, fw ? dW[1] + wOff : nullptr, fw ? dW[2] + wOff : nullptr);
	if (st != GMG_OK) return fail(st);
	if (level == 0) lap("level 0 coefficient records");
	if ((st = buildChunks(ctx, L)) != GMG_OK) return fail(st);
	if ((st = buildBricks(ctx, L)) != GMG_OK) return fail(st);
	if (level > 0 && s->lv[level - 1].bricks && (st = buildCoarseBricks(ctx, L)) != GMG_OK) return fail(st);
	if (s->opt.use_gauss_seidel && (st = buildGsTiles(ctx, L)) != GMG_OK) return fail(st);
	maxGrid = std::max(maxGrid, L.nChunksActive + int(divUp(L.nBoundary, BLOCK)) + 1);
	if ((st = allocGrid(&L.xAlt, L.g)) != GMG_OK) return fail(st);
	if ((st = allocGrid(&L.r, L.g)) != GMG_OK) return fail(st);
	if (level > 0)
	{
	    if ((st = allocGrid(&L.x, L.g)) != GMG_OK) return fail(st);
	    if ((st = allocGrid(&L.b, L.g)) != GMG_OK) return fail(st);
	}
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(cudaFail(cudaGetLastError(), "cudaStreamSynchronize", __FILE__, __LINE__));
    for (int a = 0; a < 3; ++a) devFree(dW[a]);
    lap("bands, records, chunks, grids");
    if ((st = buildIoGroups(s)) != GMG_OK) return fail(st);
    lap("transfer plan");
    if ((st = ensureScratch(ctx, maxGrid)) != GMG_OK) return fail(st);
    if (!s->opt.operators_only && (st = buildCoarseSolve(s)) != GMG_OK) return fail(st);
    if ((st = buildFusedCycle(s)) != GMG_OK) return fail(st);
    if ((st = buildClusterSmooth(s)) != GMG_OK) return fail(st);
    lap("coarse direct solver");
    if ((st = p2pEnsure(s)) != GMG_OK) return fail(st);
    if (s->shardLevels > 0)
    {
	// run every communication pattern once now: NCCL sets up its peer connections on first use, which must not
	// happen inside the stream capture of the V-cycle / PCG graphs
	for (int l = 0; l < s->shardLevels; ++l)
	    if ((st = haloExchange(s, l, s->lv[l].r, 1)) != GMG_OK) return fail(st);
	if ((st = gatherReplicated(s, s->lv[s->shardLevels].b)) != GMG_OK) return fail(st);
	if ((st = allreduceScalar(s, scalarPtr(s, offsetof(Scalars, tmp)))) != GMG_OK) return fail(st);
	if ((st = allreduceScalar(s, scalarPtr(s, offsetof(Scalars, tmp)), NCCL_MAX)) != GMG_OK) return fail(st);
	lap("communication warm-up");
    }
    if ((st = finishCoefsSparse(ctx, s->lv[0], &coefJob)) != GMG_OK) return fail(st);
    lap("level 0 coefficient records (join)");
    // (after the join: the tiles copy the diagonals and coefficient codes of level 0)
    if ((st = buildBandTiles(s)) != GMG_OK) return fail(st);
    lap("band tiles");
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(cudaFail(cudaGetLastError(), "cudaStreamSynchronize", __FILE__, __LINE__));
    s->setupMs = nowMs() - tStart;
    *out = s;
    return GMG_OK;
}

extern "C" int gmg_solver_create(gmg_ctx *ctx, const int32_t *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2,
				 int mgLevels, const gmg_solver_options *optIn, gmg_solver **out)
{
    return solverCreate(ctx, labels, res, w0, w1, w2, mgLevels, optIn, out);
}
// the same constructor on ONE-BYTE labels (the enum of Ops.h:11 as uint8): a quarter of the PCIe and host-memory traffic
extern "C" int gmg_solver_create_u8(gmg_ctx *ctx, const uint8_t *labels, const int64_t res[3], const double *w0, const double *w1, const double *w2,
				    int mgLevels, const gmg_solver_options *optIn, gmg_solver **out)
{
    return solverCreate(ctx, labels, res, w0, w1, w2, mgLevels, optIn, out);
}

extern "C" int gmg_solver_levels(gmg_solver *s, int *levels)
{
    if (!s || !levels) return invalid("null argument");
    *levels = s->levels;
    return GMG_OK;
}
extern "C" int gmg_solver_level_res(gmg_solver *s, int level, int64_t res[3])
{
    if (!s || level < 0 || level >= s->levels) return invalid("level out of range");
    for (int a = 0; a < 3; ++a) res[a] = s->lv[level].g.res[a];
    return GMG_OK;
}
extern "C" int gmg_solver_get_labels(gmg_solver *s, int level, int32_t *out)
{
    if (!s || !out || level < 0 || level >= s->levels) return invalid("level out of range");
    GMG_CUDA(enterCtx(s->ctx));
    const Level &L = s->lv[level];
    return downloadLabels(s->ctx, out, L.labelsAlloc ? L.labelsAlloc : L.labels, L.gg.res, L.gg, true);
}
extern "C" int gmg_solver_get_boundary_cells(gmg_solver *s, int level, int64_t *xyz, int64_t *count)
{
    if (!s || !count || level < 0 || level >= s->levels) return invalid("level out of range");
    GMG_CUDA(enterCtx(s->ctx));
    return exportBand(s->ctx, s->lv[level], xyz, count);
}
extern "C" int gmg_solver_active_cells(gmg_solver *s, int level, int64_t *count)
{
    if (!s || !count || level < 0 || level >= s->levels) return invalid("level out of range");
    *count = s->lv[level].nActiveGlobal;
    return GMG_OK;
}
extern "C" int gmg_solver_shard_info(gmg_solver *s, int level, int *sharded, int64_t *ownLoZ, int64_t *ownHiZ, int64_t *localActive)
{
    if (!s || level < 0 || level >= s->levels) return invalid("level out of range");
    const Level &L = s->lv[level];
    if (sharded) *sharded = L.sharded ? 1 : 0;
    if (ownLoZ) *ownLoZ = int64_t(L.g.org[2]) + L.ownLo;  // expanded z coordinates [lo, hi) of the rank's owned planes
    if (ownHiZ) *ownHiZ = int64_t(L.g.org[2]) + L.ownHi;
    if (localActive) *localActive = L.nActive;
    return GMG_OK;
}
extern "C" int gmg_solver_coarse_unknowns(gmg_solver *s, int64_t *count)
{
    if (!s || !count) return invalid("null argument");
    *count = s->nCoarse;
    return GMG_OK;
}
extern "C" int gmg_solver_transfer_cells(gmg_solver *s, int64_t *cells, int64_t *copies)
{
    if (!s || !cells) return invalid("null argument");
    *cells = s->ioCells;
    if (copies) *copies = s->ioGroups.empty() ? 1 : int64_t(s->ioGroups.size());
    return GMG_OK;
}

extern "C" int gmg_solver_setup_ms(gmg_solver *s, double *ms)
{
    if (!s || !ms) return invalid("null argument");
    *ms = s->setupMs;
    return GMG_OK;
}

// ====================================================================================================
// operator launches
// ====================================================================================================
static StencilArgs stencilArgs(gmg_solver *s, int level, const double *in, const double *b, double *out)
{
    const Level &L = s->lv[level];
    StencilArgs a;
    a.labels = L.labels;
    a.flags = L.nbrMask;
    a.in = in;
    a.b = b;
    a.out = out;
    a.chunks = L.chunksInterior;
    a.nChunks = L.nChunksInterior;
    a.chunksPerPlane = L.g.chunksPerPlane;
    a.pitch = L.g.pitch;
    a.plane = L.g.plane;
    a.nz = L.g.n[2];
    a.zlo = 0;
    a.zhi = L.g.n[2];
    a.dotLo = L.ownLo;
    a.dotHi = L.ownHi;
    a.nBoundary = L.nBoundary;
    a.bandIdx = L.bandIdx;
    a.bcoef = L.bcoef;
    a.wcode = L.wcode;
    a.partials = s->ctx->partials;
    a.ticket = s->ctx->ticket;
    a.result = nullptr;
    return a;
}

static int launchStencil(gmg_solver *s, int level, int mode, const double *in, const double *b, double *out, double *dotResult, ZRange zr)
{
    s->ctx->curLevel = level;
    const Level &L = s->lv[level];
    StencilArgs a = stencilArgs(s, level, in, b, out);
    a.result = dotResult;
    a.zlo = zr.lo;
    a.zhi = zr.hi;
    unsigned grid = unsigned(L.nChunksInterior + divUp(L.nBoundary, BLOCK));
    if (grid == 0 && !dotResult) return GMG_OK;
    grid = std::max(grid, 1u);  // a fused dot product on a slab without active cells still delivers its 0 (one idle CTA)
    cudaStream_t st = s->ctx->stream;
    const double n = double(L.nActive);
    if (L.bricks && mode != SM_JACOBI_ZERO && (s->tmaMask & (mode == SM_JACOBI ? 1 : (mode == SM_RESIDUAL ? 2 : 4))))
    {
	// TMA-staged variant: one CTA per 64 x 8 x 4 brick
	TmaMap tm;
	GMG_TRY(tensorMapOf(s, level, in, &tm));
	const unsigned tgrid = std::max(1u, unsigned(L.nBricks + divUp(L.nBoundary, BLOCK)));
	const int ny = L.g.n[1];
	if (mode == SM_JACOBI)
	{
	    GMG_LAUNCH(s->ctx, KC_JACOBI, n * 25.0);
	    GMG_CUDA(launchK((k_stencil_tma<SM_JACOBI, false>), tgrid, unsigned(BLOCK), size_t(0), st, a, tm, L.bricks, L.nBricks, L.bricksX, L.bricksY, ny));
	}
	else if (mode == SM_RESIDUAL)
	{
	    GMG_LAUNCH(s->ctx, KC_RESIDUAL, n * 25.0);
	    GMG_CUDA(launchK((k_stencil_tma<SM_RESIDUAL, false>), tgrid, unsigned(BLOCK), size_t(0), st, a, tm, L.bricks, L.nBricks, L.bricksX, L.bricksY, ny));
	}
	else if (dotResult)
	{
	    GMG_LAUNCH(s->ctx, KC_APPLY, n * 17.0);
	    GMG_CUDA(launchK((k_stencil_tma<SM_APPLY, true>), tgrid, unsigned(BLOCK), size_t(0), st, a, tm, L.bricks, L.nBricks, L.bricksX, L.bricksY, ny));
	}
	else
	{
	    GMG_LAUNCH(s->ctx, KC_APPLY, n * 17.0);
	    GMG_CUDA(launchK((k_stencil_tma<SM_APPLY, false>), tgrid, unsigned(BLOCK), size_t(0), st, a, tm, L.bricks, L.nBricks, L.bricksX, L.bricksY, ny));
	}
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    // stencilCap: bit 0 Jacobi, bit 1 residual, bit 2 apply, bit 3 zero-aware Jacobi -- the 40-register instantiation (6 resident CTAs per SM instead of 4);
    // bit 4 zero-aware Jacobi, bit 5 Jacobi at 48 registers (5 CTAs per SM)
    if (s->stencilLoop)
    {
	// persistent variant: what fits the device at once, every CTA walking its chunks with the next chunk's labels prefetched
	gmg_ctx *ctx = s->ctx;
	if (ctx->stencilSlots == 0)
	{
	    int per = 0;
	    GMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_stencil_loop<SM_JACOBI, false>, BLOCK, 0));
	    ctx->stencilSlots = std::max(1, per) * ctx->smCount;
	    if (const char *e = getenv("GMG_STENCIL_LOOP_SLOTS")) ctx->stencilSlots = std::max(1, atoi(e));  // tests: force the loop on small grids
	}
	const unsigned g = std::min(grid, unsigned(ctx->stencilSlots));
	const int bit = mode == SM_JACOBI ? 1 : (mode == SM_JACOBI_ZERO ? 8 : (mode == SM_RESIDUAL ? 2 : 4));
	if ((s->stencilLoop & bit) && grid > g)
	{
	    if (mode == SM_JACOBI)
	    {
		GMG_LAUNCH(s->ctx, KC_JACOBI, n * 25.0);
		GMG_CUDA(launchK((k_stencil_loop<SM_JACOBI, false>), g, unsigned(BLOCK), size_t(0), st, a, int(grid)));
	    }
	    else if (mode == SM_JACOBI_ZERO)
	    {
		GMG_LAUNCH(s->ctx, KC_JACOBI, n * 18.0);
		GMG_CUDA(launchK((k_stencil_loop<SM_JACOBI_ZERO, false>), g, unsigned(BLOCK), size_t(0), st, a, int(grid)));
	    }
	    else if (mode == SM_RESIDUAL)
	    {
		GMG_LAUNCH(s->ctx, KC_RESIDUAL, n * 25.0);
		GMG_CUDA(launchK((k_stencil_loop<SM_RESIDUAL, false>), g, unsigned(BLOCK), size_t(0), st, a, int(grid)));
	    }
	    else if (dotResult)
	    {
		GMG_LAUNCH(s->ctx, KC_APPLY, n * 17.0);
		GMG_CUDA(launchK((k_stencil_loop<SM_APPLY, true>), g, unsigned(BLOCK), size_t(0), st, a, int(grid)));
	    }
	    else
	    {
		GMG_LAUNCH(s->ctx, KC_APPLY, n * 17.0);
		GMG_CUDA(launchK((k_stencil_loop<SM_APPLY, false>), g, unsigned(BLOCK), size_t(0), st, a, int(grid)));
	    }
	    GMG_CUDA(cudaGetLastError());
	    return GMG_OK;
	}
    }
    if (s->stencilBatch == 2 || s->stencilBatch == 4)
    {
	// the loads of 2 / 4 planes in flight together (k_stencil_b)
#define GMG_SB(M, D)                                                                                                                   \
    do                                                                                                                                 \
    {                                                                                                                                  \
	if (s->stencilBatch == 4) GMG_CUDA(launchK((k_stencil_b<M, D, 4>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));         \
	else GMG_CUDA(launchK((k_stencil_b<M, D, 2>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));                              \
    } while (0)
	if (mode == SM_JACOBI) { GMG_LAUNCH(s->ctx, KC_JACOBI, n * 25.0); GMG_SB(SM_JACOBI, false); }
	else if (mode == SM_JACOBI_ZERO) { GMG_LAUNCH(s->ctx, KC_JACOBI, n * 18.0); GMG_SB(SM_JACOBI_ZERO, false); }
	else if (mode == SM_RESIDUAL) { GMG_LAUNCH(s->ctx, KC_RESIDUAL, n * 25.0); GMG_SB(SM_RESIDUAL, false); }
	else if (dotResult) { GMG_LAUNCH(s->ctx, KC_APPLY, n * 17.0); GMG_SB(SM_APPLY, true); }
	else { GMG_LAUNCH(s->ctx, KC_APPLY, n * 17.0); GMG_SB(SM_APPLY, false); }
#undef GMG_SB
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    // measured at 256^3 (4.3 M cells, plain loads): apply 38.3 -> 32.3 us, residual 34.6 -> 31.5 us, Jacobi no gain (it spills);
    // on a 0.5 M-cell level everything loses 5-10 % -- so: residual + apply on levels of at least 2 M cells unless GMG_STENCIL_CAP says otherwise
    const int cap = s->stencilCap >= 0 ? s->stencilCap : (n >= 2.0e6 ? 6 : 0);
    if (mode == SM_JACOBI)
    {
	GMG_LAUNCH(s->ctx, KC_JACOBI, n * 25.0);
	if (cap & 1) GMG_CUDA(launchK((k_stencil<SM_JACOBI, false, double, 6>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
	else if (cap & 32) GMG_CUDA(launchK((k_stencil<SM_JACOBI, false, double, 5>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
	else GMG_CUDA(launchK((k_stencil<SM_JACOBI, false>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
    }
    else if (mode == SM_JACOBI_ZERO)
    {
	// x is zero off the band: its 8 bytes per cell are not read (one flag byte is), and no zero fill ran before
	GMG_LAUNCH(s->ctx, KC_JACOBI, n * 18.0);
	if (cap & 8) GMG_CUDA(launchK((k_stencil<SM_JACOBI_ZERO, false, double, 6>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
	else if (cap & 16) GMG_CUDA(launchK((k_stencil<SM_JACOBI_ZERO, false, double, 5>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
	else GMG_CUDA(launchK((k_stencil<SM_JACOBI_ZERO, false>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
    }
    else if (mode == SM_RESIDUAL)
    {
	GMG_LAUNCH(s->ctx, KC_RESIDUAL, n * 25.0);
	if (cap & 2) GMG_CUDA(launchK((k_stencil<SM_RESIDUAL, false, double, 6>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
	else GMG_CUDA(launchK((k_stencil<SM_RESIDUAL, false>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
    }
    else if (dotResult)
    {
	GMG_LAUNCH(s->ctx, KC_APPLY, n * 17.0);
	if (cap & 4) GMG_CUDA(launchK((k_stencil<SM_APPLY, true, double, 6>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
	else GMG_CUDA(launchK((k_stencil<SM_APPLY, true>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
    }
    else
    {
	GMG_LAUNCH(s->ctx, KC_APPLY, n * 17.0);
	if (cap & 4) GMG_CUDA(launchK((k_stencil<SM_APPLY, false, double, 6>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
	else GMG_CUDA(launchK((k_stencil<SM_APPLY, false>), unsigned(grid), unsigned(BLOCK), size_t(0), st, a));
    }
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

// `sweeps` boundary-band Jacobi sweeps on grid x (Ops.h:524-619); zeroGrid: x is to be taken as all zero -- the grid is
// never read (it may hold anything), only its band cells are written by the last sweep
static int launchBand(gmg_solver *s, int level, double *x, const double *b, int sweeps, bool zeroGrid)
{
    s->ctx->curLevel = level;
    const Level &L = s->lv[level];
    if (L.nBand == 0 || sweeps <= 0) return GMG_OK;
    BandArgs a;
    a.x = x;
    a.b = b;
    a.bandIdx = L.bandIdx;
    a.bandRef = L.bandRef;
    a.bcoef = L.bcoef;
    a.wcode = L.wcode;
    a.bandB = L.bandB;
    a.nBoundary = L.nBoundary;
    a.nBand = L.nBand;
    a.pitch = L.g.pitch;
    a.plane = L.g.plane;
    const unsigned grid = unsigned(divUp(L.nBand, BLOCK * BAND_PER_THREAD));
    cudaStream_t st = s->ctx->stream;
    const double bytes = double(L.nBand) * 29.0;
    const bool hw = L.hasWeights;
    double *cur = L.bandV0, *nxt = L.bandV1;
    if (sweeps == 3 && L.tileArgs && s->bandTiles && (zeroGrid || L.tilesCoResident))
    {
	// the whole group as ring-halo tiles: one launch, no barrier between the sweeps (gmg_band_tiles.cuh)
	BandTileArgs t = *static_cast<const BandTileArgs *>(L.tileArgs);
	t.x = x;
	t.b = b;
	GMG_LAUNCH(s->ctx, KC_BAND, bytes * sweeps);
	if (zeroGrid) GMG_CUDA(launchK(k_band_tile<true>, unsigned(L.nTiles), unsigned(BT_THREADS), L.tileSmem, st, t));
	else GMG_CUDA(launchK(k_band_tile<false>, unsigned(L.nTiles), unsigned(BT_THREADS), L.tileSmem, st, t));
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    const bool groupHere = s->bandGroupMax <= 0 || L.nBand <= s->bandGroupMax;
    if (sweeps >= 2 && s->bandResident && groupHere && s->ctx->groupBarrier)
    {
	// the whole group in one launch of 2 x 512 threads per SM with every cell's metadata resident (k_band_resident);
	// bands beyond the capacity of 6 cells per thread take the sweep-per-launch kernels below
	gmg_ctx *ctx = s->ctx;
	const int64_t gthreads = int64_t(2) * ctx->smCount * BG_THREADS;
	const int cpt = int(divUp(L.nBand, gthreads));
	if (cpt <= 6)
	{
	    // small bands: fewer CTAs (one cell per thread), a cheaper barrier
	    const unsigned g = unsigned(std::min<int64_t>(int64_t(2) * ctx->smCount, divUp(L.nBand, BG_THREADS)));
	    const int c2 = int(divUp(L.nBand, int64_t(g) * BG_THREADS));
	    const int CP = c2 <= 1 ? 1 : (c2 <= 2 ? 2 : (c2 <= 4 ? 4 : 6));
	    const size_t smem = size_t(CP) * 6 * BG_THREADS * sizeof(int);
	    a.vin = nxt;   // (the kernel takes the two compact arrays through vin / vout)
	    a.vout = cur;
	    GroupBarrier *bar = static_cast<GroupBarrier *>(ctx->groupBarrier);
	    static bool attr = false;
	    if (!attr)
	    {
		const int big = 6 * 6 * BG_THREADS * int(sizeof(int));
		GMG_CUDA(cudaFuncSetAttribute(k_band_resident<false, false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
		GMG_CUDA(cudaFuncSetAttribute(k_band_resident<false, true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
		GMG_CUDA(cudaFuncSetAttribute(k_band_resident<true, false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
		GMG_CUDA(cudaFuncSetAttribute(k_band_resident<true, true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
		attr = true;
	    }
	    GMG_LAUNCH(ctx, KC_BAND, bytes * sweeps);
#define GMG_BR(HW, ZG)                                                                                                         \
    do                                                                                                                         \
    {                                                                                                                          \
	if (CP == 1) GMG_CUDA(launchK((k_band_resident<HW, ZG, 1>), g, unsigned(BG_THREADS), smem, st, a, sweeps, bar));       \
	else if (CP == 2) GMG_CUDA(launchK((k_band_resident<HW, ZG, 2>), g, unsigned(BG_THREADS), smem, st, a, sweeps, bar));  \
	else if (CP == 4) GMG_CUDA(launchK((k_band_resident<HW, ZG, 4>), g, unsigned(BG_THREADS), smem, st, a, sweeps, bar));  \
	else GMG_CUDA(launchK((k_band_resident<HW, ZG, 6>), g, unsigned(BG_THREADS), smem, st, a, sweeps, bar));               \
    } while (0)
	    if (zeroGrid && hw) GMG_BR(true, true);
	    else if (zeroGrid) GMG_BR(false, true);
	    else if (hw) GMG_BR(true, false);
	    else GMG_BR(false, false);
#undef GMG_BR
	    GMG_CUDA(cudaGetLastError());
	    return GMG_OK;
	}
    }
    if (sweeps >= 2 && s->bandGroups && groupHere && s->ctx->groupBarrier)
    {
	// the whole group in one co-resident launch with grid barriers between the sweeps (k_band_group)
	gmg_ctx *ctx = s->ctx;
	if (ctx->bandGroupCtas[0] == 0)
	{
	    int per = 0;
	    GMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_band_group<false, false>, BLOCK, 0));
	    ctx->bandGroupCtas[0] = std::max(1, per) * ctx->smCount;
	    GMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_band_group<true, false>, BLOCK, 0));
	    ctx->bandGroupCtas[1] = std::max(1, per) * ctx->smCount;
	}
	const unsigned cap = unsigned(ctx->bandGroupCtas[hw && !zeroGrid ? 1 : 0]);
	const unsigned g = std::min(grid, cap);
	a.vin = nxt;   // (the kernel takes the two compact arrays through vin / vout)
	a.vout = cur;
	GroupBarrier *bar = static_cast<GroupBarrier *>(ctx->groupBarrier);
	GMG_LAUNCH(ctx, KC_BAND, bytes * sweeps);
	if (zeroGrid && hw) GMG_CUDA(launchK((k_band_group<true, true>), g, BLOCK, 0, st, a, sweeps, int(grid), bar));
	else if (zeroGrid) GMG_CUDA(launchK((k_band_group<false, true>), g, BLOCK, 0, st, a, sweeps, int(grid), bar));
	else if (hw) GMG_CUDA(launchK((k_band_group<true, false>), g, BLOCK, 0, st, a, sweeps, int(grid), bar));
	else GMG_CUDA(launchK((k_band_group<false, false>), g, BLOCK, 0, st, a, sweeps, int(grid), bar));
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    // cells per thread: two, or three where two would need a second wave of CTAs and three do not
    gmg_ctx *ctx = s->ctx;
    if (ctx->bandSlots[0] == 0)
    {
	int per = 0;
	GMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_band<true, false, false, false, true, false, double, 2>, BLOCK, 0));
	ctx->bandSlots[0] = std::max(1, per) * ctx->smCount;
	GMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_band<true, false, false, false, true, false, double, 3>, BLOCK, 0));
	ctx->bandSlots[1] = std::max(1, per) * ctx->smCount;
    }
    const unsigned grid3 = unsigned(divUp(L.nBand, BLOCK * 3));
    // measured (profiles/r02_band_fusion.md): no gain at 256^3 level 0 (the one level where it applies), a loss when forced everywhere:
    // two cells per thread unless GMG_BAND_PER_THREAD=3 / =-1 (pick per level) asks otherwise
    bool pt3 = s->bandPerThread == -1 && int(grid) > ctx->bandSlots[0] && int(grid3) <= ctx->bandSlots[1];
    if (s->bandPerThread == 2) pt3 = false;
    if (s->bandPerThread == 3) pt3 = true;
    const unsigned gridK = pt3 ? grid3 : grid;
#define GMG_BAND(FC, TG, FI, ZE, HW, FZ)                                                                                            \
    do                                                                                                                             \
    {                                                                                                                              \
	if (pt3) GMG_CUDA(launchK((k_band<FC, TG, FI, ZE, HW, FZ, double, 3>), gridK, BLOCK, 0, st, a));                            \
	else GMG_CUDA(launchK((k_band<FC, TG, FI, ZE, HW, FZ, double, 2>), gridK, BLOCK, 0, st, a));                                \
    } while (0)
    // profiling mode 2: one bracket around the whole group (the launches' own brackets below are muted inside it)
    LaunchScope groupBracket(s->ctx, KC_BAND, bytes * sweeps, sweeps, true);
    // sweep 1: grid -> compact
    a.vin = nullptr;
    a.vout = cur;
    {
	GMG_LAUNCH(s->ctx, KC_BAND, bytes);
	if (zeroGrid) GMG_BAND(false, false, true, true, false, false);
	else if (hw) GMG_BAND(false, false, true, false, true, false);
	else GMG_BAND(false, false, true, false, false, false);
    }
    if (sweeps == 1)
    {
	GMG_LAUNCH(s->ctx, KC_BAND, double(L.nBand) * 20.0);
	GMG_CUDA(launchK(k_band_scatter<double>, unsigned(divUp(L.nBand, BLOCK)), BLOCK, 0, st, x, L.bandIdx, cur, L.nBand));
    }
    for (int sw = 2; sw <= sweeps; ++sw)
    {
	a.vin = cur;
	a.vout = nxt;
	GMG_LAUNCH(s->ctx, KC_BAND, bytes);
	if (sw == sweeps)
	{
	    if (zeroGrid && hw) GMG_BAND(true, true, false, false, true, true);
	    else if (zeroGrid) GMG_BAND(true, true, false, false, false, true);
	    else if (hw) GMG_BAND(true, true, false, false, true, false);
	    else GMG_BAND(true, true, false, false, false, false);
	}
	else
	{
	    if (zeroGrid && hw) GMG_BAND(true, false, false, false, true, true);
	    else if (zeroGrid) GMG_BAND(true, false, false, false, false, true);
	    else if (hw) GMG_BAND(true, false, false, false, true, false);
	    else GMG_BAND(true, false, false, false, false, false);
	}
	std::swap(cur, nxt);
    }
#undef GMG_BAND
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

static TransferArgs transferArgs(gmg_solver *s, int fineLevel)
{
    const Level &F = s->lv[fineLevel], &C = s->lv[fineLevel + 1];
    TransferArgs a;
    a.fineLabels = F.labels;
    a.coarseLabels = C.labels;
    a.finePitch = F.g.pitch;
    a.coarsePitch = C.g.pitch;
    a.finePlane = F.g.plane;
    a.coarsePlane = C.g.plane;
    a.fineNz = F.g.n[2];
    a.coarseNz = C.g.n[2];
    a.coarseNy = C.g.n[1];
    for (int k = 0; k < 3; ++k) a.shift[k] = F.shift[k];
    a.fine = nullptr; a.coarse = nullptr; a.out = nullptr; a.chunks = nullptr; a.chunksPerPlane = 0;
    a.zlo = 0; a.zhi = 0;
    return a;
}

static int launchRestrict(gmg_solver *s, int fineLevel, double *coarse, const double *fine, ZRange zr)
{
    s->ctx->curLevel = fineLevel + 1;
    const Level &C = s->lv[fineLevel + 1];
    if (C.nChunksActive == 0) return GMG_OK;
    TransferArgs a = transferArgs(s, fineLevel);
    a.fine = fine;
    a.out = coarse;
    a.chunks = C.chunksActive;
    a.chunksPerPlane = C.g.chunksPerPlane;
    a.zlo = zr.lo;
    a.zhi = zr.hi;
    GMG_LAUNCH(s->ctx, KC_RESTRICT, double(s->lv[fineLevel].nActive) * 8.0 + double(C.nActive) * 9.0);
    if (C.cbricks && (s->tmaMask & 8))
    {
	// TMA-staged variant: one CTA per 32 x 4 x 2 brick of coarse cells
	TmaMap tm;
	GMG_TRY(tensorMapOf(s, fineLevel, fine, &tm));
	if (C.nCBricks > 0) GMG_CUDA(launchK(k_restrict_tma, unsigned(C.nCBricks), unsigned(BLOCK), size_t(0), s->ctx->stream, a, tm, C.cbricks, C.cbricksX, C.cbricksY));
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    GMG_CUDA(launchK(k_restrict<double>, unsigned(C.nChunksActive * RESTRICT_SPLIT), unsigned(BLOCK), size_t(0), s->ctx->stream, a));
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

static int launchProlong(gmg_solver *s, int fineLevel, double *fine, const double *coarse, ZRange zr)
{
    s->ctx->curLevel = fineLevel;
    const Level &F = s->lv[fineLevel];
    if (F.nChunksActive == 0) return GMG_OK;
    TransferArgs a = transferArgs(s, fineLevel);
    a.coarse = coarse;
    a.out = fine;
    a.chunks = F.chunksActive;
    a.chunksPerPlane = F.g.chunksPerPlane;
    a.zlo = zr.lo;
    a.zhi = zr.hi;
    GMG_LAUNCH(s->ctx, KC_PROLONG, double(F.nActive) * 17.0 + double(s->lv[fineLevel + 1].nActive) * 8.0);
    if (F.bricksActive && (s->tmaMask & 16))
    {
	// TMA-staged variant: the coarse box of every 64 x 8 x 4 fine brick through shared memory
	TmaMap tm;
	GMG_TRY(tensorMapOf(s, fineLevel + 1, coarse, &tm, 1));
	if (F.nBricksActive > 0)
	    GMG_CUDA(launchK(k_prolong_tma, unsigned(F.nBricksActive), unsigned(BLOCK), size_t(0), s->ctx->stream, a, tm, F.bricksActive, F.bricksX, F.bricksY, F.g.n[1]));
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    GMG_CUDA(launchK(k_prolong<double>, unsigned(F.nChunksActive), unsigned(BLOCK), size_t(0), s->ctx->stream, a));
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

static int launchZero(gmg_solver *s, int level, double *x)
{
    s->ctx->curLevel = level;
    const Level &L = s->lv[level];
    if (L.nChunksActive == 0) return GMG_OK;
    GMG_LAUNCH(s->ctx, KC_ZERO, double(L.nActive) * 8.0);
    GMG_CUDA(launchK(k_zero, unsigned(L.nChunksActive), unsigned(BLOCK), size_t(0), s->ctx->stream, x, L.chunksActive, L.g.chunksPerPlane, L.g.plane, L.g.n[2]));
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

static int launchCoarse(gmg_solver *s, double *x, const double *b)
{
    s->ctx->curLevel = s->levels - 1;
    const int n = s->nCoarse;
    GMG_LAUNCH(s->ctx, KC_COARSE, double(n) * n * 8.0);
    GMG_CUDA(launchK(k_coarse_solve, unsigned(unsigned(divUp(n, BLOCK / 32))), unsigned(BLOCK), size_t(sizeof(double) * n), s->ctx->stream, x, b, s->coarseIdx, s->coarseInv, n));
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

// levels [fusedFirst, levels-1] in ONE thread-block cluster with the vectors in distributed shared memory
// (gmg_cluster.cuh: k_cluster_cycle).  Tables are built on the device, stream-ordered, without a host synchronisation.
static int buildClusterCycle(gmg_solver *s)
{
    s->fusedFirst = -1;
    // OPT-IN (GMG_CLUSTER_CYCLE=1).  Measured on B200 (profiles/r02_cluster_cycle.md): 195 us for levels 3..6 of the 256^3
    // problem against 98 us for level 3 as kernels of its own plus the one-CTA cycle for levels 4..6 -- half of it the
    // single-CTA levels (every restriction tap / prolongation corner between a distributed level and a single-CTA level funnels
    // through one SM's ~20 B/cycle cluster port), a quarter the cluster barriers themselves (0.5 us each, scripts/cluster_probe.cu).
    {
	const char *e = getenv("GMG_CLUSTER_CYCLE");
	if (!(e && e[0] == '1')) return GMG_OK;
    }
    if (const char *e = getenv("GMG_COARSE_FUSED")) if (e[0] == '0') return GMG_OK;
    if (s->opt.operators_only || s->levels < 2) return GMG_OK;
    if (s->opt.use_gauss_seidel) return GMG_OK;  // the fused cycles implement the Jacobi interior sweep only
    gmg_ctx *ctx = s->ctx;
    if (ctx->clusterSize < 0) return GMG_OK;     // probed before: this device / driver refuses the cluster launch
    int smemMax = 0;
    GMG_CUDA(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    if (ctx->clusterSize == 0)
    {
	// one probe per context: the largest cluster (16 needs the non-portable opt-in) that can be resident with the maximum
	// dynamic shared memory
	ctx->clusterSize = -1;
	if (cudaFuncSetAttribute(k_cluster_cycle, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax) != cudaSuccess) { cudaGetLastError(); return GMG_OK; }
	if (cudaFuncSetAttribute(k_cluster_cycle, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
	int want = 16;
	if (const char *e = getenv("GMG_CLUSTER_SIZE")) want = std::max(1, std::min(16, atoi(e)));
	for (int cl = want; cl >= 2; cl >>= 1)
	{
	    cudaLaunchConfig_t cfg = {};
	    cfg.gridDim = dim3(cl);
	    cfg.blockDim = dim3(CLUSTER_THREADS);
	    cfg.dynamicSmemBytes = size_t(smemMax);
	    cudaLaunchAttribute attr[1];
	    attr[0].id = cudaLaunchAttributeClusterDimension;
	    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	    cfg.attrs = attr;
	    cfg.numAttrs = 1;
	    int nClusters = 0;
	    if (cudaOccupancyMaxActiveClusters(&nClusters, k_cluster_cycle, &cfg) == cudaSuccess && nClusters >= 1) { ctx->clusterSize = cl; break; }
	    cudaGetLastError();
	}
	if (ctx->clusterSize < 0) return GMG_OK;
    }
    const int CL = ctx->clusterSize;
    // cells per CTA block: a level of at most CLUSTER_SOLO_MAX cells lives in CTA 0 alone (its steps end on __syncthreads, not on
    // the cluster barrier); larger levels get an equal share per CTA, but never less than one cell per thread
    auto perOf = [&](int64_t n) {
	n = std::max<int64_t>(n, 1);
	const int64_t per = n <= CLUSTER_SOLO_MAX ? n : std::max<int64_t>(divUp(n, CL), std::min<int64_t>(n, CLUSTER_THREADS));
	return int((per + 1) & ~int64_t(1));
    };
    // per cell of a block: two solution arrays + rhs (24 bytes) and the smoother tables (6 x 16-bit codes, diagonal, flags: 14 bytes)
    auto smemOf = [&](int first) {
	size_t bytes = 0;
	for (int l = first; l < s->levels; ++l) bytes += size_t(38) * perOf(s->lv[l].nActive) + 32;
	return bytes + (size_t(s->nCoarse) + 2) * sizeof(double);
    };
    // finest level (>= 1: level 0 carries face weights and the caller's grids; replicated levels only) whose vectors fit the
    // cluster's shared memory with at most 8 cells per thread
    int first = -1;
    for (int l = std::max(1, s->shardLevels); l < s->levels; ++l)
	if (s->levels - l <= CLUSTER_MAX_LEVELS && (s->lv[l].nActive <= CLUSTER_SOLO_MAX || perOf(s->lv[l].nActive) <= CLUSTER_MAX_PER) &&
	    smemOf(l) + 1024 <= size_t(smemMax))
	{
	    first = l;
	    break;
	}
    if (const char *c = getenv("GMG_FUSED_FIRST")) first = std::max(first, atoi(c));
    if (first < 0 || first >= s->levels - 1) return GMG_OK;  // nothing, or only the direct solve: its own kernel does that
    const int nl = s->levels - first;
    // one slab for every table
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    struct Offs { size_t cell, nbr, nbr16, rst, pro, diag, flags; };
    std::vector<Offs> o(nl);
    for (int q = 0; q < nl; ++q)
    {
	const size_t n = size_t(std::max<int64_t>(s->lv[first + q].nActive, 1));
	o[q].cell = take(4 * n);
	o[q].nbr = take(4 * 6 * n);
	o[q].nbr16 = take(2 * 6 * n);
	o[q].rst = q > 0 ? take(4 * 64 * n) : 0;
	o[q].pro = q + 1 < nl ? take(4 * 8 * n) : 0;
	o[q].diag = take(n);
	o[q].flags = take(n);
    }
    const size_t oSolve = take(4 * size_t(std::max(s->nCoarse, 1)));
    char *slab = nullptr;
    GMG_CUDA(devMalloc(&slab, off));
    ClusterArgs *c = new ClusterArgs;
    std::memset(c, 0, sizeof(*c));
    s->clusterArgs = c;
    s->clusterSlab = slab;
    // compact lists and position grids (storage index -> compact index), level by level
    std::vector<int32_t *> pos(nl, nullptr);
    int smemOff = 0;
    for (int q = 0; q < nl; ++q)
    {
	const Level &L = s->lv[first + q];
	const int n = int(L.nActive);
	ClusterLevel &K = c->lv[q];
	K.n = n;
	K.per = perOf(n);
	K.off = smemOff;
	smemOff += 3 * K.per;
	K.nbr = reinterpret_cast<const unsigned *>(slab + o[q].nbr);
	K.nbr16 = reinterpret_cast<const unsigned short *>(slab + o[q].nbr16);
	K.rst = reinterpret_cast<const unsigned *>(slab + o[q].rst);
	K.pro = reinterpret_cast<const unsigned *>(slab + o[q].pro);
	K.diag = reinterpret_cast<const uint8_t *>(slab + o[q].diag);
	K.flags = reinterpret_cast<const uint8_t *>(slab + o[q].flags);
	int32_t *cell = reinterpret_cast<int32_t *>(slab + o[q].cell);
	const int64_t total = L.g.total;
	uint8_t *fl = nullptr;
	int *dCount = nullptr;
	void *dTemp = nullptr;
	size_t tempBytes = 0;
	GMG_CUDA(devMalloc(&fl, size_t(total)));
	GMG_CUDA(devMalloc(&dCount, sizeof(int)));
	GMG_CUDA(devMalloc(&pos[q], sizeof(int32_t) * size_t(total)));
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_active_flags<<<unsigned(divUp(total, BLOCK)), BLOCK, 0, ctx->stream>>>(fl, L.labels, total);
	}
	thrust::counting_iterator<int32_t> it(0);
	GMG_CUDA(cub::DeviceSelect::Flagged(dTemp, tempBytes, it, fl, cell, dCount, int(total), ctx->stream));
	GMG_CUDA(devMalloc(&dTemp, std::max<size_t>(tempBytes, 16)));
	GMG_CUDA(cub::DeviceSelect::Flagged(dTemp, tempBytes, it, fl, cell, dCount, int(total), ctx->stream));
	++ctx->launches;
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_fill_i32<<<unsigned(divUp(total, BLOCK)), BLOCK, 0, ctx->stream>>>(pos[q], -1, total);
	}
	if (n > 0)
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_band_pos<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(pos[q], cell, n);
	}
	GMG_CUDA(devFree(dTemp));
	GMG_CUDA(devFree(dCount));
	GMG_CUDA(devFree(fl));
    }
    for (int q = 0; q < nl; ++q)
    {
	const Level &L = s->lv[first + q];
	const Geom &g = L.g;
	const ClusterLevel &K = c->lv[q];
	const int n = K.n;
	if (n == 0) continue;
	const int32_t *cell = reinterpret_cast<const int32_t *>(slab + o[q].cell);
	const unsigned grid = unsigned(divUp(n, BLOCK));
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_cluster_nbr<<<grid, BLOCK, 0, ctx->stream>>>(const_cast<unsigned *>(K.nbr), const_cast<uint8_t *>(K.diag), const_cast<uint8_t *>(K.flags), cell, pos[q],
							   L.labels, L.bandFlags, n, K.per, g.pitch, g.plane);
	    k_cluster_nbr16<<<grid, BLOCK, 0, ctx->stream>>>(const_cast<unsigned short *>(K.nbr16), K.nbr, n, K.per);
	}
	if (q > 0)
	{
	    const Level &F = s->lv[first + q - 1];
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_cluster_rst<<<grid, BLOCK, 0, ctx->stream>>>(const_cast<unsigned *>(K.rst), cell, pos[q - 1], n, c->lv[q - 1].per, g.pitch, g.plane, F.g.pitch, F.g.plane,
							   F.g.n[0], F.g.n[1], F.g.n[2], F.shift[0], F.shift[1], F.shift[2]);
	}
	if (q + 1 < nl)
	{
	    const Level &C = s->lv[first + q + 1];
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_cluster_pro<<<grid, BLOCK, 0, ctx->stream>>>(const_cast<unsigned *>(K.pro), cell, pos[q + 1], n, c->lv[q + 1].per, g.pitch, g.plane, C.g.pitch, C.g.plane,
							   C.g.n[0], C.g.n[1], C.g.n[2], L.shift[0], L.shift[1], L.shift[2]);
	}
    }
    unsigned *solveRef = reinterpret_cast<unsigned *>(slab + oSolve);
    if (s->nCoarse > 0)
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_cluster_solve_ref<<<unsigned(divUp(s->nCoarse, BLOCK)), BLOCK, 0, ctx->stream>>>(solveRef, s->coarseIdx, pos[nl - 1], s->nCoarse, c->lv[nl - 1].per);
    }
    GMG_CUDA(cudaGetLastError());
    for (int q = 0; q < nl; ++q) GMG_CUDA(devFree(pos[q]));
    c->nLevels = nl;
    c->soloFirst = nl;
    for (int q = nl - 1; q >= 0 && s->lv[first + q].nActive <= CLUSTER_SOLO_MAX; --q) c->soloFirst = q;
    c->sweeps = s->opt.boundary_iterations;
    c->scratchOff = smemOff;
    {
	// table region behind the vectors and the direct solve's scratch (16-byte aligned pieces)
	size_t tab = (size_t(smemOff + s->nCoarse + 2) * sizeof(double) + 15) & ~size_t(15);
	for (int q = 0; q < nl; ++q)
	{
	    c->lv[q].tabOff = int(tab);
	    tab = (tab + size_t(14) * c->lv[q].per + 15) & ~size_t(15);
	}
	s->clusterSmem = tab;
    }
    c->cellTop = reinterpret_cast<const int32_t *>(slab + o[0].cell);
    c->bTop = s->lv[first].b;
    c->xTop = s->lv[first].x;
    c->solveRef = solveRef;
    c->nSolve = s->nCoarse;
    c->inv = s->coarseInv;
    if (s->clusterSmem > size_t(smemMax)) return invalid("cluster cycle: shared-memory estimate too tight");
    s->fusedFirst = first;
    return GMG_OK;
}

// one level's smoothing in a cluster (gmg_cluster.cuh: k_cluster_smooth): levels of 2k..16k cells above the fused coarse cycle
static int buildClusterSmooth(gmg_solver *s)
{
    {
	const char *e = getenv("GMG_CLUSTER_SMOOTH");
	if (e && e[0] == '0') return GMG_OK;
    }
    if (s->opt.operators_only || s->opt.use_gauss_seidel || s->opt.boundary_iterations < 1) return GMG_OK;
    gmg_ctx *ctx = s->ctx;
    int smemMax = 0;
    GMG_CUDA(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    if (ctx->smoothClusterSize == 0)
    {
	ctx->smoothClusterSize = -1;
	if (cudaFuncSetAttribute(k_cluster_smooth, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax) != cudaSuccess) { cudaGetLastError(); return GMG_OK; }
	if (cudaFuncSetAttribute(k_cluster_smooth, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
	for (int cl = 16; cl >= 8; cl >>= 1)
	{
	    cudaLaunchConfig_t cfg = {};
	    cfg.gridDim = dim3(cl);
	    cfg.blockDim = dim3(CLUSTER_THREADS);
	    cfg.dynamicSmemBytes = 64 << 10;
	    cudaLaunchAttribute attr[1];
	    attr[0].id = cudaLaunchAttributeClusterDimension;
	    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	    cfg.attrs = attr;
	    cfg.numAttrs = 1;
	    int nClusters = 0;
	    if (cudaOccupancyMaxActiveClusters(&nClusters, k_cluster_smooth, &cfg) == cudaSuccess && nClusters >= 1) { ctx->smoothClusterSize = cl; break; }
	    cudaGetLastError();
	}
    }
    if (ctx->smoothClusterSize < 0) return GMG_OK;
    const int CL = ctx->smoothClusterSize;
    const int last = s->fusedFirst > 0 ? s->fusedFirst : s->levels - 1;  // levels [1, last) run as kernels
    for (int level = std::max(1, s->shardLevels); level < last; ++level)
    {
	Level &L = s->lv[level];
	const int n = int(L.nActive);
	if (L.nActive < 2048 || L.nActive > int64_t(CL) * CLUSTER_THREADS * 2) continue;
	const int per = int((divUp(n, CL) + 1) & ~int64_t(1));
	if (per > CLUSTER_MAX_PER) continue;
	size_t off = 0;
	auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
	const size_t oCell = take(4 * size_t(n)), oNbr = take(4 * 6 * size_t(n)), oNbr16 = take(2 * 6 * size_t(n)), oDiag = take(n), oFlags = take(n);
	char *slab = nullptr;
	GMG_CUDA(devMalloc(&slab, off));
	ClusterSmoothArgs *c = new ClusterSmoothArgs;
	std::memset(c, 0, sizeof(*c));
	L.smoothArgs = c;
	L.smoothSlab = slab;
	ClusterLevel &K = c->lv;
	K.n = n; K.per = per; K.off = 0;
	K.tabOff = int((size_t(3) * per * sizeof(double) + 15) & ~size_t(15));
	K.nbr = reinterpret_cast<const unsigned *>(slab + oNbr);
	K.nbr16 = reinterpret_cast<const unsigned short *>(slab + oNbr16);
	K.diag = reinterpret_cast<const uint8_t *>(slab + oDiag);
	K.flags = reinterpret_cast<const uint8_t *>(slab + oFlags);
	L.smoothSmem = size_t(K.tabOff) + size_t(14) * per + 16;
	int32_t *cell = reinterpret_cast<int32_t *>(slab + oCell);
	const int64_t total = L.g.total;
	uint8_t *fl = nullptr;
	int *dCount = nullptr;
	int32_t *pos = nullptr;
	void *dTemp = nullptr;
	size_t tempBytes = 0;
	GMG_CUDA(devMalloc(&fl, size_t(total)));
	GMG_CUDA(devMalloc(&dCount, sizeof(int)));
	GMG_CUDA(devMalloc(&pos, sizeof(int32_t) * size_t(total)));
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_active_flags<<<unsigned(divUp(total, BLOCK)), BLOCK, 0, ctx->stream>>>(fl, L.labels, total);
	}
	thrust::counting_iterator<int32_t> it(0);
	GMG_CUDA(cub::DeviceSelect::Flagged(dTemp, tempBytes, it, fl, cell, dCount, int(total), ctx->stream));
	GMG_CUDA(devMalloc(&dTemp, std::max<size_t>(tempBytes, 16)));
	GMG_CUDA(cub::DeviceSelect::Flagged(dTemp, tempBytes, it, fl, cell, dCount, int(total), ctx->stream));
	++ctx->launches;
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_fill_i32<<<unsigned(divUp(total, BLOCK)), BLOCK, 0, ctx->stream>>>(pos, -1, total);
	    k_band_pos<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(pos, cell, n);
	    k_cluster_nbr<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(const_cast<unsigned *>(K.nbr), const_cast<uint8_t *>(K.diag), const_cast<uint8_t *>(K.flags), cell,
										pos, L.labels, L.bandFlags, n, per, L.g.pitch, L.g.plane);
	    k_cluster_nbr16<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(const_cast<unsigned short *>(K.nbr16), K.nbr, n, per);
	}
	GMG_CUDA(cudaGetLastError());
	GMG_CUDA(devFree(dTemp));
	GMG_CUDA(devFree(dCount));
	GMG_CUDA(devFree(fl));
	GMG_CUDA(devFree(pos));
	c->sweeps = s->opt.boundary_iterations;
	c->cell = cell;
    }
    return GMG_OK;
}

// band sweeps + interior sweep + band sweeps (+ residual on the way down) of one level as ONE cluster kernel
static int launchClusterSmooth(gmg_solver *s, int level, double *x, const double *b, double *r, bool up)
{
    gmg_ctx *ctx = s->ctx;
    ctx->curLevel = level;
    Level &L = s->lv[level];
    ClusterSmoothArgs a = *static_cast<const ClusterSmoothArgs *>(L.smoothArgs);
    a.x = x; a.b = b; a.r = r;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctx->smoothClusterSize);
    cfg.blockDim = dim3(CLUSTER_THREADS);
    cfg.dynamicSmemBytes = L.smoothSmem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctx->smoothClusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = usePdl() ? 2 : 1;
    GMG_LAUNCH(ctx, KC_BAND, double(L.nBand) * 29.0 * 2 * s->opt.boundary_iterations + double(L.nActive) * (up ? 25.0 : 50.0));
    GMG_CUDA(cudaLaunchKernelEx(&cfg, k_cluster_smooth, a, up ? 1 : 0));
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

// ------------------------------------------------------------------------------------------------
// Ring-halo tiles of a level's band (gmg_band_tiles.cuh): built once per solver, on the device.
// ------------------------------------------------------------------------------------------------
static int sortKeys64(gmg_ctx *ctx, unsigned long long *keys, unsigned long long *out, int64_t n, int bits)
{
    void *dTemp = nullptr;
    size_t tempBytes = 0;
    GMG_CUDA(cub::DeviceRadixSort::SortKeys(dTemp, tempBytes, keys, out, n, 0, bits, ctx->stream));
    GMG_CUDA(devMalloc(&dTemp, std::max<size_t>(tempBytes, 16)));
    GMG_CUDA(cub::DeviceRadixSort::SortKeys(dTemp, tempBytes, keys, out, n, 0, bits, ctx->stream));
    ++ctx->launches;
    GMG_CUDA(devFree(dTemp));
    return GMG_OK;
}
// sorted keys -> the first of every run of equal (key >> shift), empty keys dropped; *out is allocated here
static int uniqueKeys64(gmg_ctx *ctx, const unsigned long long *sorted, int64_t n, int shift, unsigned long long **out, int *count)
{
    uint8_t *head = nullptr;
    int *dCount = nullptr;
    unsigned long long *tmp = nullptr;
    GMG_CUDA(devMalloc(&head, size_t(std::max<int64_t>(n, 1))));
    GMG_CUDA(devMalloc(&dCount, sizeof(int)));
    GMG_CUDA(devMalloc(&tmp, sizeof(unsigned long long) * size_t(std::max<int64_t>(n, 1))));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_tile_heads<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(head, sorted, n, shift);
    }
    void *dTemp = nullptr;
    size_t tempBytes = 0;
    GMG_CUDA(cub::DeviceSelect::Flagged(dTemp, tempBytes, sorted, head, tmp, dCount, int(n), ctx->stream));
    GMG_CUDA(devMalloc(&dTemp, std::max<size_t>(tempBytes, 16)));
    GMG_CUDA(cub::DeviceSelect::Flagged(dTemp, tempBytes, sorted, head, tmp, dCount, int(n), ctx->stream));
    ++ctx->launches;
    int h = 0;
    GMG_CUDA(cudaMemcpyAsync(&h, dCount, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = h;
    *out = tmp;
    GMG_CUDA(devFree(dTemp));
    GMG_CUDA(devFree(dCount));
    GMG_CUDA(devFree(head));
    return GMG_OK;
}

static int buildLevelTiles(gmg_solver *s, Level &L)
{
    gmg_ctx *ctx = s->ctx;
    const int nBand = L.nBand;
    const int own = int(std::min<int64_t>(1536, std::max<int64_t>(256, divUp(divUp(nBand, int64_t(2) * ctx->smCount), 128) * 128)));
    const int nTiles = int(divUp(nBand, own));
    const unsigned gridBand = unsigned(divUp(nBand, BLOCK));
    cudaStream_t st = ctx->stream;
    // 1. Morton order of the band cells, equal cuts
    unsigned *key = nullptr, *keyOut = nullptr;
    int32_t *val = nullptr, *order = nullptr, *tileOf = nullptr, *sortedPos = nullptr;
    GMG_CUDA(devMalloc(&key, sizeof(unsigned) * nBand));
    GMG_CUDA(devMalloc(&keyOut, sizeof(unsigned) * nBand));
    GMG_CUDA(devMalloc(&val, sizeof(int32_t) * nBand));
    GMG_CUDA(devMalloc(&order, sizeof(int32_t) * nBand));
    GMG_CUDA(devMalloc(&tileOf, sizeof(int32_t) * nBand));
    GMG_CUDA(devMalloc(&sortedPos, sizeof(int32_t) * nBand));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_tile_keys<<<gridBand, BLOCK, 0, st>>>(key, val, L.bandIdx, nBand, L.g.pitch, L.g.plane);
    }
    {
	void *dTemp = nullptr;
	size_t tempBytes = 0;
	GMG_CUDA(cub::DeviceRadixSort::SortPairs(dTemp, tempBytes, key, keyOut, val, order, nBand, 0, 30, st));
	GMG_CUDA(devMalloc(&dTemp, std::max<size_t>(tempBytes, 16)));
	GMG_CUDA(cub::DeviceRadixSort::SortPairs(dTemp, tempBytes, key, keyOut, val, order, nBand, 0, 30, st));
	++ctx->launches;
	GMG_CUDA(devFree(dTemp));
    }
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_tile_assign<<<gridBand, BLOCK, 0, st>>>(tileOf, sortedPos, order, nBand, own);
    }
    GMG_CUDA(devFree(key)); GMG_CUDA(devFree(keyOut)); GMG_CUDA(devFree(val));
    // 2. ring 1: band neighbours in another tile
    unsigned long long *k1 = nullptr, *k1s = nullptr, *ring1 = nullptr;
    int n1 = 0;
    GMG_CUDA(devMalloc(&k1, sizeof(unsigned long long) * 6 * size_t(nBand)));
    GMG_CUDA(devMalloc(&k1s, sizeof(unsigned long long) * 6 * size_t(nBand)));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_tile_ring1<<<gridBand, BLOCK, 0, st>>>(k1, L.bandRef, tileOf, nBand);
    }
    GMG_TRY(sortKeys64(ctx, k1, k1s, int64_t(6) * nBand, 64));
    GMG_TRY(uniqueKeys64(ctx, k1s, int64_t(6) * nBand, 0, &ring1, &n1));
    GMG_CUDA(devFree(k1)); GMG_CUDA(devFree(k1s));
    // 3. ring 2: band neighbours of ring 1 outside the tile and ring 1; 4. the halo of a tile sorted (ring, position)
    unsigned long long *halo = nullptr;
    int nHalo = 0;
    int *dCnt = nullptr;
    GMG_CUDA(devMalloc(&dCnt, sizeof(int) * 2 * nTiles));
    GMG_CUDA(cudaMemsetAsync(dCnt, 0, sizeof(int) * 2 * nTiles, st));
    if (n1 > 0)
    {
	unsigned long long *k2 = nullptr, *k2s = nullptr, *uni = nullptr;
	GMG_CUDA(devMalloc(&k2, sizeof(unsigned long long) * 7 * size_t(n1)));
	GMG_CUDA(devMalloc(&k2s, sizeof(unsigned long long) * 7 * size_t(n1)));
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_tile_ring2<<<unsigned(divUp(n1, BLOCK)), BLOCK, 0, st>>>(k2, ring1, n1, L.bandRef, tileOf, nBand);
	}
	GMG_TRY(sortKeys64(ctx, k2, k2s, int64_t(7) * n1, 64));
	GMG_TRY(uniqueKeys64(ctx, k2s, int64_t(7) * n1, 1, &uni, &nHalo));
	GMG_CUDA(devFree(k2)); GMG_CUDA(devFree(k2s));
	{
	    GMG_LAUNCH(ctx, KC_SETUP, 0);
	    k_tile_rekey<<<unsigned(divUp(nHalo, BLOCK)), BLOCK, 0, st>>>(uni, dCnt, nHalo);
	}
	GMG_CUDA(devMalloc(&halo, sizeof(unsigned long long) * size_t(nHalo)));
	GMG_TRY(sortKeys64(ctx, uni, halo, nHalo, 64));
	GMG_CUDA(devFree(uni));
    }
    GMG_CUDA(devFree(ring1));
    // 5. tile headers
    std::vector<int> cnt(2 * size_t(nTiles));
    GMG_CUDA(cudaMemcpyAsync(cnt.data(), dCnt, sizeof(int) * 2 * nTiles, cudaMemcpyDeviceToHost, st));
    GMG_CUDA(cudaStreamSynchronize(st));
    GMG_CUDA(devFree(dCnt));
    std::vector<int4> tiles(nTiles);
    std::vector<int32_t> haloStart(nTiles);
    int maxLoc = 0, maxCalc = 0;
    int64_t totalLoc = 0, hs = 0, sumR1 = 0, sumR2 = 0;
    for (int t = 0; t < nTiles; ++t)
    {
	const int nOwn = std::min(own, nBand - t * own);
	const int r1 = cnt[2 * t], r2 = cnt[2 * t + 1];
	tiles[t] = make_int4(int(totalLoc), nOwn, nOwn + r1, nOwn + r1 + r2);
	haloStart[t] = int32_t(hs);
	maxLoc = std::max(maxLoc, nOwn + r1 + r2);
	maxCalc = std::max(maxCalc, nOwn + r1);
	totalLoc += nOwn + r1 + r2;
	hs += r1 + r2;
	sumR1 += r1;
	sumR2 += r2;
    }
    auto release = [&]() {
	devFree(order); devFree(tileOf); devFree(sortedPos); devFree(halo);
    };
    const size_t smem = size_t(14) * maxLoc + size_t(28) * maxCalc;
    int smemMax = 0;
    GMG_CUDA(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    if (maxLoc > BT_MAX_LOC || smem > size_t(smemMax) || totalLoc > int64_t(0x7fffffff) / 8) { release(); return GMG_OK; }
    // 6. per-tile tables
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    const size_t oTiles = take(sizeof(int4) * nTiles), oHs = take(sizeof(int32_t) * nTiles), oGi = take(sizeof(int32_t) * totalLoc), oJ = take(sizeof(int32_t) * totalLoc),
		 oCode = take(sizeof(unsigned short) * totalLoc), oDiag = take(sizeof(double) * totalLoc), oRef = take(sizeof(unsigned short) * 6 * totalLoc), oErr = take(sizeof(int));
    char *slab = nullptr;
    GMG_CUDA(devMalloc(&slab, off));
    L.tileSlab = slab;
    GMG_CUDA(cudaMemcpyAsync(slab + oTiles, tiles.data(), sizeof(int4) * nTiles, cudaMemcpyHostToDevice, st));
    GMG_CUDA(cudaMemcpyAsync(slab + oHs, haloStart.data(), sizeof(int32_t) * nTiles, cudaMemcpyHostToDevice, st));
    GMG_CUDA(cudaMemsetAsync(slab + oErr, 0, sizeof(int), st));
    TileFillArgs f;
    f.tiles = reinterpret_cast<const int4 *>(slab + oTiles);
    f.haloStart = reinterpret_cast<const int32_t *>(slab + oHs);
    f.halo = halo;
    f.order = order; f.tileOf = tileOf; f.sortedPos = sortedPos; f.bandIdx = L.bandIdx; f.bandRef = L.bandRef;
    f.bcoef = L.bcoef; f.wcode = L.wcode;
    f.locGi = reinterpret_cast<int32_t *>(slab + oGi);
    f.locJ = reinterpret_cast<int32_t *>(slab + oJ);
    f.locCode = reinterpret_cast<unsigned short *>(slab + oCode);
    f.locRef = reinterpret_cast<unsigned short *>(slab + oRef);
    f.locDiag = reinterpret_cast<double *>(slab + oDiag);
    f.error = reinterpret_cast<int *>(slab + oErr);
    f.nTiles = nTiles; f.own = own; f.nBand = nBand; f.nBoundary = L.nBoundary; f.hasWeights = L.hasWeights ? 1 : 0; f.totalLoc = int(totalLoc);
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_tile_fill<<<unsigned(divUp(totalLoc, BLOCK)), BLOCK, 0, st>>>(f);
    }
    int err = 0;
    GMG_CUDA(cudaMemcpyAsync(&err, slab + oErr, sizeof(int), cudaMemcpyDeviceToHost, st));
    GMG_CUDA(cudaStreamSynchronize(st));
    GMG_CUDA(cudaGetLastError());
    release();
    if (err)
    {
	// a neighbour of a computed cell is missing from the tile: never expected; keep the sweep-per-launch kernels
	devFree(L.tileSlab);
	L.tileSlab = nullptr;
	return GMG_OK;
    }
    BandTileArgs *a = new BandTileArgs;
    std::memset(a, 0, sizeof(*a));
    a->tiles = f.tiles; a->locGi = f.locGi; a->locJ = f.locJ; a->locCode = f.locCode; a->locDiag = f.locDiag; a->locRef = f.locRef;
    a->bcoef = L.bcoef;
    a->bar = static_cast<GroupBarrier *>(ctx->groupBarrier);
    a->nBoundary = std::max(L.nBoundary, 1);
    a->pitch = L.g.pitch; a->plane = L.g.plane;
    a->maxLoc = maxLoc; a->maxCalc = maxCalc;
    L.tileArgs = a;
    L.tileSmem = smem;
    L.nTiles = nTiles;
    static bool attr = false;
    if (!attr)
    {
	GMG_CUDA(cudaFuncSetAttribute(k_band_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax));
	GMG_CUDA(cudaFuncSetAttribute(k_band_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax));
	attr = true;
    }
    int per = 0;
    GMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_band_tile<false>, BT_THREADS, smem));
    L.tilesCoResident = int64_t(per) * ctx->smCount >= nTiles;
    if (const char *e = getenv("GMG_BAND_TILES_ZERO_ONLY")) { if (e[0] == '1') L.tilesCoResident = false; }
    if (s->opt.print_stats)
	printf("      band tiles: %d cells -> %d tiles of %d (+%.0f %% ring 1, +%.0f %% ring 2), %zu B shared memory, %d CTAs per SM%s\n", nBand, nTiles, own,
	       100.0 * double(sumR1) / nBand, 100.0 * double(sumR2) / nBand, smem, per, L.tilesCoResident ? "" : " (not co-resident: zero-grid groups only)");
    return GMG_OK;
}

static int buildBandTiles(gmg_solver *s)
{
    if (!s->bandTiles || s->opt.operators_only || s->opt.boundary_iterations != 3) return GMG_OK;
    const int last = s->fusedFirst > 0 ? s->fusedFirst : s->levels - 1;  // levels [0, last) run as kernels
    for (int level = 0; level < last; ++level)
    {
	Level &L = s->lv[level];
	if (L.smoothArgs || L.nBand < 512) continue;
	GMG_TRY(buildLevelTiles(s, L));
    }
    return GMG_OK;
}

// levels [fusedFirst, levels-1] in one shared-memory CTA (gmg_kernels.cuh: k_compact_cycle): the fallback where the cluster
// cycle is not available.  The tables are built on the host: these levels hold a few thousand cells at most.
static int buildFusedCycle(gmg_solver *s)
{
    GMG_TRY(buildClusterCycle(s));
    if (s->fusedFirst > 0) return GMG_OK;
    s->fusedFirst = -1;
    const char *e = getenv("GMG_COARSE_FUSED");
    if (e && e[0] == '0') return GMG_OK;
    if (s->opt.operators_only || s->levels < 2) return GMG_OK;
    if (s->opt.use_gauss_seidel) return GMG_OK;  // the compact cycle implements the Jacobi interior sweep only
    gmg_ctx *ctx = s->ctx;
    int smemMax = 0;
    GMG_CUDA(cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    // finest level (>= 1: level 0 carries face weights and the caller's grids; replicated levels only) whose vectors
    // (3 doubles per cell) and 16-bit index tables (<= 12 + 128 + 16 + 2 bytes per cell) fit one CTA's shared memory
    auto smemOf = [&](int first) {
	size_t bytes = 0;
	for (int l = first; l < s->levels; ++l)
	{
	    const size_t n = size_t(s->lv[l].nActive);
	    bytes += n * (3 * sizeof(double) + 12 + 2 + 16) + (l > first ? n * 128 : 0) + 64;
	}
	return bytes + size_t(s->nCoarse) * 2 + 64;
    };
    int first = -1;
    for (int l = std::max(1, s->shardLevels); l < s->levels; ++l)
	if (s->levels - l <= CYCLE_MAX_LEVELS && s->lv[l].nActive < 65535 && smemOf(l) + 1024 <= size_t(smemMax)) { first = l; break; }
    if (const char *c = getenv("GMG_FUSED_FIRST")) first = std::max(first, atoi(c));
    if (first < 0 || first >= s->levels) return GMG_OK;
    if (first == s->levels - 1) return GMG_OK;  // only the direct solve: its own kernel does that
    const int nl = s->levels - first;
    struct HostLevel
    {
	std::vector<int32_t> cell, pos;
	std::vector<uint16_t> nbr, rst, pro;
	std::vector<uint8_t> diag, flags, lab;
    };
    std::vector<HostLevel> h(nl);
    for (int q = 0; q < nl; ++q)
    {
	const Level &L = s->lv[first + q];
	const Geom &g = L.g;
	HostLevel &H = h[q];
	H.lab.resize(g.total);
	GMG_CUDA(cudaMemcpy(H.lab.data(), L.labels, g.total, cudaMemcpyDeviceToHost));
	H.pos.assign(g.total, -1);
	for (int64_t i = 0; i < g.total; ++i)
	    if (H.lab[i] == L_INTERIOR || H.lab[i] == L_BOUNDARY) { H.pos[i] = int32_t(H.cell.size()); H.cell.push_back(int32_t(i)); }
	const int n = int(H.cell.size());
	if (n != L.nActive) return invalid("compact coarse cycle: active-cell count mismatch");
	H.nbr.assign(size_t(6) * n, CYCLE_NONE);
	H.diag.assign(n, 0);
	H.flags.assign(n, 0);
	const int64_t stride[6] = {-1, 1, -int64_t(g.pitch), int64_t(g.pitch), -g.plane, g.plane};
	for (int k = 0; k < n; ++k)
	{
	    const int64_t i = H.cell[k];
	    int diag = 0;
	    for (int d = 0; d < 6; ++d)
	    {
		const int nlab = H.lab[i + stride[d]];
		if (nlab == L_INTERIOR || nlab == L_BOUNDARY) { H.nbr[size_t(d) * n + k] = uint16_t(H.pos[i + stride[d]]); ++diag; }
		else if (nlab == L_DIRICHLET) ++diag;
	    }
	    H.diag[k] = uint8_t(diag);  // 6 for an INTERIOR cell (all neighbours active by construction)
	    const int z = int(i / g.plane);
	    const int64_t rem = i - int64_t(z) * g.plane;
	    const int y = int(rem / g.pitch), x = int(rem - int64_t(y) * g.pitch);
	    H.flags[k] = uint8_t(((x & 1) << 1) | ((y & 1) << 2) | ((z & 1) << 3));
	}
	std::vector<int32_t> band(std::max(L.nBand, 1));
	GMG_CUDA(cudaMemcpy(band.data(), L.bandIdx, sizeof(int32_t) * L.nBand, cudaMemcpyDeviceToHost));
	for (int k = 0; k < L.nBand; ++k) H.flags[H.pos[band[k]]] |= 1;
    }
    auto compact = [](int32_t p) { return p < 0 ? uint16_t(CYCLE_NONE) : uint16_t(p); };
    for (int q = 0; q < nl; ++q)
    {
	const Level &L = s->lv[first + q];
	const Geom &g = L.g;
	HostLevel &H = h[q];
	const int n = int(H.cell.size());
	if (q > 0)
	{
	    // restriction taps of this (coarse) level in the next finer compact level (Ops.h:760-834)
	    const Level &F = s->lv[first + q - 1];
	    const HostLevel &HF = h[q - 1];
	    H.rst.assign(size_t(64) * n, CYCLE_NONE);
	    for (int k = 0; k < n; ++k)
	    {
		const int64_t i = H.cell[k];
		const int cz = int(i / g.plane);
		const int64_t rem = i - int64_t(cz) * g.plane;
		const int cy = int(rem / g.pitch), cx = int(rem - int64_t(cy) * g.pitch);
		const int fx = 2 * (cx - F.shift[0]) - 1, fy = 2 * (cy - F.shift[1]) - 1, fz = 2 * (cz - F.shift[2]) - 1;
		for (int z = 0; z < 4; ++z)
		    for (int y = 0; y < 4; ++y)
			for (int x = 0; x < 4; ++x)
			{
			    const int X = fx + x, Y = fy + y, Z = fz + z;
			    if (X < 0 || Y < 0 || Z < 0 || X >= F.g.n[0] || Y >= F.g.n[1] || Z >= F.g.n[2]) continue;
			    H.rst[size_t((z * 4 + y) * 4 + x) * n + k] = compact(HF.pos[int64_t(Z) * F.g.plane + int64_t(Y) * F.g.pitch + X]);
			}
	    }
	}
	if (q + 1 < nl)
	{
	    // prolongation corners of this (fine) level in the next coarser compact level (Ops.h:895-971)
	    const Level &C = s->lv[first + q + 1];
	    const HostLevel &HC = h[q + 1];
	    H.pro.assign(size_t(8) * n, CYCLE_NONE);
	    for (int k = 0; k < n; ++k)
	    {
		const int64_t i = H.cell[k];
		const int fz = int(i / g.plane);
		const int64_t rem = i - int64_t(fz) * g.plane;
		const int fy = int(rem / g.pitch), fx = int(rem - int64_t(fy) * g.pitch);
		const int mx = (fx >> 1) + L.shift[0], my = (fy >> 1) + L.shift[1], mz = (fz >> 1) + L.shift[2];
		const int xs = (fx & 1) ? mx : mx - 1, ys = (fy & 1) ? my : my - 1, zs = (fz & 1) ? mz : mz - 1;
		for (int c = 0; c < 8; ++c)
		{
		    const int X = xs + (c & 1), Y = ys + ((c >> 1) & 1), Z = zs + (c >> 2);
		    if (X < 0 || Y < 0 || Z < 0 || X >= C.g.n[0] || Y >= C.g.n[1] || Z >= C.g.n[2]) continue;
		    H.pro[size_t(c) * n + k] = compact(HC.pos[int64_t(Z) * C.g.plane + int64_t(Y) * C.g.pitch + X]);
		}
	    }
	}
    }
    // direct-solve numbering -> compact index of the last level
    std::vector<uint16_t> solveIdx(s->nCoarse);
    std::vector<int32_t> coarseIdx(s->nCoarse);
    GMG_CUDA(cudaMemcpy(coarseIdx.data(), s->coarseIdx, sizeof(int32_t) * s->nCoarse, cudaMemcpyDeviceToHost));
    for (int k = 0; k < s->nCoarse; ++k) solveIdx[k] = compact(h[nl - 1].pos[coarseIdx[k]]);
    // table blob (goes to shared memory as a whole; 16-byte aligned pieces), followed by the top level's cell list
    std::vector<char> blob;
    auto put = [&](const void *p, size_t bytes) {
	const size_t off = (blob.size() + 15) & ~size_t(15);
	blob.resize(off + bytes);
	if (bytes) std::memcpy(blob.data() + off, p, bytes);
	return int(off);
    };
    CompactArgs *c = new CompactArgs;
    std::memset(c, 0, sizeof(*c));
    int off = 0;
    for (int q = 0; q < nl; ++q)
    {
	HostLevel &H = h[q];
	CompactLevel &L = c->lv[q];
	L.n = int(H.cell.size());
	L.off = off;
	off += 3 * L.n;
	L.nbr = put(H.nbr.data(), H.nbr.size() * 2);
	L.rst = put(H.rst.data(), H.rst.size() * 2);
	L.pro = put(H.pro.data(), H.pro.size() * 2);
	L.diag = put(H.diag.data(), H.diag.size());
	L.flags = put(H.flags.data(), H.flags.size());
    }
    c->solveIdx = put(solveIdx.data(), solveIdx.size() * 2);
    blob.resize((blob.size() + 15) & ~size_t(15));
    c->blobBytes = int(blob.size());
    c->vectorDoubles = (off + 1) & ~1;
    const size_t cellOff = blob.size();
    blob.resize(cellOff + h[0].cell.size() * 4);
    std::memcpy(blob.data() + cellOff, h[0].cell.data(), h[0].cell.size() * 4);
    const size_t smem = size_t(c->vectorDoubles) * sizeof(double) + size_t(c->blobBytes);
    if (smem > size_t(smemMax)) { delete c; return GMG_OK; }  // estimate was too tight: keep the per-kernel path
    char *d = nullptr;
    GMG_CUDA(devMalloc(&d, blob.size()));
    GMG_CUDA(cudaMemcpyAsync(d, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    c->blob = reinterpret_cast<const unsigned char *>(d);
    c->cellTop = reinterpret_cast<const int32_t *>(d + cellOff);
    c->nLevels = nl;
    c->sweeps = s->opt.boundary_iterations;
    c->bTop = s->lv[first].b;
    c->xTop = s->lv[first].x;
    c->inv = s->coarseInv;
    c->nSolve = s->nCoarse;
    s->compactArgs = c;
    s->compactBlob = d;
    s->compactSmem = smem;
    // the attribute is per function, not per solver: always the device maximum, so solvers of different sizes can coexist
    GMG_CUDA(cudaFuncSetAttribute(k_compact_cycle, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax));
    s->fusedFirst = first;
    return GMG_OK;
}

static int launchCoarseCycle(gmg_solver *s)
{
    s->ctx->curLevel = s->fusedFirst;
    double bytes = 0;
    for (int l = s->fusedFirst; l < s->levels - 1; ++l) bytes += double(s->lv[l].nActive) * 126.0;
    GMG_LAUNCH(s->ctx, KC_COARSE, bytes);
    if (s->clusterArgs)
    {
	const ClusterArgs &cc = *static_cast<const ClusterArgs *>(s->clusterArgs);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(s->ctx->clusterSize);
	cfg.blockDim = dim3(CLUSTER_THREADS);
	cfg.dynamicSmemBytes = s->clusterSmem;
	cfg.stream = s->ctx->stream;
	cudaLaunchAttribute attr[2];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = s->ctx->clusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[1].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = usePdl() ? 2 : 1;
	GMG_CUDA(cudaLaunchKernelEx(&cfg, k_cluster_cycle, cc));
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    const CompactArgs &c = *static_cast<const CompactArgs *>(s->compactArgs);
    GMG_CUDA(launchK(k_compact_cycle, unsigned(1), unsigned(CYCLE_THREADS), size_t(s->compactSmem), s->ctx->stream, c));
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

template <int OP>
static int launchVec(gmg_solver *s, int level, double *y, const double *a, const double *c, double *y2, double sc, double *result, int klass,
		     double bytesPerCell, ZRange zr = {0, 1 << 30})
{
    s->ctx->curLevel = level;
    const Level &L = s->lv[level];
    // a slab without active cells (a sharded level whose liquid does not reach this rank) has nothing to update, but a
    // REDUCTION still has to deliver its 0 -- and the CG update has to retire rho -- on this rank like on every other
    const bool reduces = OP == VO_DOT || OP == VO_NORM2 || OP == VO_MAX || OP == VO_CG_UPDATE;
    if (L.nChunksActive == 0 && !reduces) return GMG_OK;
    VecArgs v;
    v.chunks = L.chunksActive;
    v.nChunks = L.nChunksActive;
    v.chunksPerPlane = L.g.chunksPerPlane;
    v.plane = L.g.plane;
    v.nz = L.g.n[2];
    v.zlo = std::max(0, zr.lo);
    v.zhi = std::min(L.g.n[2], zr.hi);
    v.redLo = L.ownLo;
    v.redHi = L.ownHi;
    v.y = y;
    v.a = a;
    v.c = c;
    v.y2 = y2;
    v.s = sc;
    v.sc = reinterpret_cast<Scalars *>(s->ctx->scalars);
    v.partials = s->ctx->partials;
    v.ticket = s->ctx->ticket;
    v.result = result;
    GMG_LAUNCH(s->ctx, klass, double(L.nActive) * bytesPerCell);
    GMG_CUDA(launchK((k_vec<OP>), unsigned(std::max(L.nChunksActive, 1)), unsigned(BLOCK), size_t(0), s->ctx->stream, v));
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

static double *scalarPtr(gmg_solver *s, size_t offset) { return reinterpret_cast<double *>(reinterpret_cast<char *>(s->ctx->scalars) + offset); }

static int readScalar(gmg_solver *s, size_t offset, double *out)
{
    GMG_CUDA(cudaMemcpyAsync(s->ctx->hostScalars, scalarPtr(s, offset), sizeof(double), cudaMemcpyDeviceToHost, s->ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(s->ctx->stream));
    *out = *s->ctx->hostScalars;
    return GMG_OK;
}

// ====================================================================================================
// V-cycle (MG.cpp:420-881)
// ====================================================================================================
// one half-pass of the tiled Gauss-Seidel smoother, in place (Ops.h:369-520)
static int launchGaussSeidel(gmg_solver *s, int level, double *x, const double *b, bool oddTiles, bool forward)
{
    s->ctx->curLevel = level;
    const Level &L = s->lv[level];
    if (!L.bpos) return invalid("this solver was not created with use_gauss_seidel");
    const int n = L.nGsTiles[oddTiles ? 1 : 0];
    if (n == 0) return GMG_OK;
    GsArgs a;
    a.x = x; a.b = b; a.labels = L.labels; a.tiles = L.gsTiles[oddTiles ? 1 : 0]; a.bpos = L.bpos; a.bcoef = L.bcoef; a.nBoundary = L.nBoundary;
    a.tilesX = L.gsTilesX; a.tilesY = L.gsTilesY;
    for (int k = 0; k < 3; ++k) { a.off[k] = L.gsOff[k]; a.n[k] = L.g.n[k]; }
    a.pitch = L.g.pitch; a.plane = L.g.plane; a.forward = forward ? 1 : 0;
    static const bool v1 = [] { const char *e = getenv("GMG_GS_V1"); return e && e[0] == '1'; }();
    if (!v1)
    {
	// records of the tile's BOUNDARY cells gathered into shared memory before the sweep (k_gauss_seidel2)
	GsArgs2 a2;
	a2.g = a;
	a2.wcode = L.wcode;
	const size_t smem2 = sizeof(double) * (GS_HALO * GS_HALO * GS_HALO + GS_TILE * GS_TILE * GS_TILE + GS_POOL) + sizeof(unsigned short) * (GS_POOL + GS_TILE * GS_TILE * GS_TILE) +
			     GS_TILE * GS_TILE * GS_TILE;
	static bool attr2 = false;
	if (!attr2) { GMG_CUDA(cudaFuncSetAttribute(k_gauss_seidel2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem2))); attr2 = true; }
	GMG_LAUNCH(s->ctx, KC_GS, double(L.nActive) * 25.0 * 0.5);
	GMG_CUDA(launchK(k_gauss_seidel2, unsigned(n), unsigned(BLOCK), size_t(smem2), s->ctx->stream, a2));
	GMG_CUDA(cudaGetLastError());
	return GMG_OK;
    }
    const size_t smem = sizeof(double) * (GS_HALO * GS_HALO * GS_HALO + GS_TILE * GS_TILE * GS_TILE) + GS_TILE * GS_TILE * GS_TILE;
    static bool attr = false;
    if (!attr) { GMG_CUDA(cudaFuncSetAttribute(k_gauss_seidel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); attr = true; }
    GMG_LAUNCH(s->ctx, KC_GS, double(L.nActive) * 25.0 * 0.5);
    GMG_CUDA(launchK(k_gauss_seidel, unsigned(n), unsigned(BLOCK), size_t(smem), s->ctx->stream, a));
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

// Zero-aware down-stroke: no zero fill of x before the first band sweeps, no read of x in the interior sweep after them
// (DESIGN.md section 4).  Needs the damped-Jacobi interior smoother and at least one band sweep (the band sweeps are what
// define the grid's band cells); GMG_ZERO_AWARE=0 restores the zero fill.
static bool zeroAware(const gmg_solver *s)
{
    return s->zeroAware && !s->opt.use_gauss_seidel && s->opt.boundary_iterations >= 1;
}

// jacobiDepth: how far into the halo the interior sweep still produces valid values (sharded levels)
// down: the smoothing before the coarse-grid correction (GS: odd then even tiles, forwards; MG.cpp:466-479), else after it
// (GS: even then odd tiles, backwards; MG.cpp:740-751)
static int smoothLevel(gmg_solver *s, int level, double *&cur, double *&alt, const double *b, bool zeroGrid, int jacobiDepth, bool down)
{
    const int it = s->opt.boundary_iterations;
    GMG_TRY(launchBand(s, level, cur, b, it, zeroGrid));
    if (s->opt.use_gauss_seidel)
    {
	GMG_TRY(launchGaussSeidel(s, level, cur, b, down, down));
	GMG_TRY(launchGaussSeidel(s, level, cur, b, !down, down));
    }
    else
    {
	// zeroGrid: the band sweeps above wrote the band cells of a grid that was NOT zero-filled; the interior sweep takes
	// every other cell as zero (SM_JACOBI_ZERO) and writes every active cell of alt
	GMG_TRY(launchStencil(s, level, zeroGrid && zeroAware(s) ? SM_JACOBI_ZERO : SM_JACOBI, cur, b, alt, nullptr, clipDepth(s->lv[level], jacobiDepth)));
	std::swap(cur, alt);
    }
    GMG_TRY(launchBand(s, level, cur, b, it, false));
    return GMG_OK;
}

// Halo bookkeeping on a sharded level (it = 3 band sweeps): the rhs is valid HALO_X = 8 planes deep.  Down: x = 0, so the
// first sweep is free and after 3 + 1 + 3 sweeps x is valid 2 deep, the residual 1 deep -- what the restriction of the owned
// coarse planes reads.  Up: the prolonged x is refreshed 8 deep, 7 sweeps leave it valid 1 deep -- what the next finer
// prolongation reads.  With an initial guess at level 0 the first sweep is not free: x and b must be valid HALO_P = 9 deep.
static int restrictDown(gmg_solver *s, int level, const double *r)
{
    Level &F = s->lv[level], &C = s->lv[level + 1];
    ZRange zr = {0, C.g.n[2]};
    if (C.sharded) zr = {C.ownLo, C.ownHi};
    else if (F.sharded) zr = {s->gatherLo[s->ctx->rank], s->gatherHi[s->ctx->rank]};
    GMG_TRY(launchRestrict(s, level, C.b, r, zr));
    if (C.sharded) GMG_TRY(haloExchange(s, level + 1, C.b, HALO_X));
    else if (F.sharded) GMG_TRY(gatherReplicated(s, C.b));
    return GMG_OK;
}

static int vcycleLaunches(gmg_solver *s, double *x, const double *b, bool useInitialGuess)
{
    const int nl = s->levels;
    const int it = s->opt.boundary_iterations;
    const int downJacobi = HALO_X - (it - 1) - 1, upJacobi = HALO_X - it - 1;
    // level 0 works on the caller's grid and the level's alternate; two Jacobi sweeps per level bring the
    // result back into the caller's buffer (one sweep only when there is a single level)
    double *cur0 = x, *alt0 = s->lv[0].xAlt;
    if (!useInitialGuess && !zeroAware(s)) GMG_TRY(launchZero(s, 0, cur0));
    GMG_TRY(smoothLevel(s, 0, cur0, alt0, b, !useInitialGuess, downJacobi, true));
    if (nl == 1)
    {
	if (cur0 != x) GMG_TRY((launchVec<VO_COPY>(s, 0, x, cur0, nullptr, nullptr, 0, nullptr, KC_BLAS1, 16.0)));
	return GMG_OK;
    }
    GMG_TRY(launchStencil(s, 0, SM_RESIDUAL, cur0, b, s->lv[0].r, nullptr, clipDepth(s->lv[0], 1)));
    GMG_TRY(restrictDown(s, 0, s->lv[0].r));
    std::vector<double *> cur(nl), alt(nl);
    // levels [fusedFirst, nl-1] run inside the persistent cluster kernel; its result is lv[fusedFirst].x
    const int nReg = (s->fusedFirst > 0) ? s->fusedFirst : nl - 1;  // regular levels are 1 .. nReg-1
    for (int level = 1; level < nReg; ++level)
    {
	Level &L = s->lv[level];
	cur[level] = L.x;
	alt[level] = L.xAlt;
	if (L.smoothArgs)
	{
	    // the level's whole down-stroke smoothing and residual in one cluster kernel (result in cur[level] = L.x, L.r)
	    GMG_TRY(launchClusterSmooth(s, level, cur[level], L.b, L.r, false));
	    GMG_TRY(restrictDown(s, level, L.r));
	    continue;
	}
	if (!zeroAware(s)) GMG_TRY(launchZero(s, level, cur[level]));
	GMG_TRY(smoothLevel(s, level, cur[level], alt[level], L.b, true, downJacobi, true));
	GMG_TRY(launchStencil(s, level, SM_RESIDUAL, cur[level], L.b, L.r, nullptr, clipDepth(L, 1)));
	GMG_TRY(restrictDown(s, level, L.r));
    }
    if (s->fusedFirst > 0) GMG_TRY(launchCoarseCycle(s));
    else GMG_TRY(launchCoarse(s, s->lv[nl - 1].x, s->lv[nl - 1].b));
    for (int level = nReg - 1; level >= 1; --level)
    {
	Level &L = s->lv[level];
	// the level below hands over its result in its own x grid (two Jacobi swaps, or the direct solve / fused cycle)
	GMG_TRY(launchProlong(s, level, cur[level], s->lv[level + 1].x, clipDepth(L, 0)));
	GMG_TRY(haloExchange(s, level, cur[level], HALO_X));
	if (L.smoothArgs) { GMG_TRY(launchClusterSmooth(s, level, cur[level], L.b, nullptr, true)); continue; }
	GMG_TRY(smoothLevel(s, level, cur[level], alt[level], L.b, false, upJacobi, false));
    }
    GMG_TRY(launchProlong(s, 0, cur0, s->lv[1].x, clipDepth(s->lv[0], 0)));
    GMG_TRY(haloExchange(s, 0, cur0, HALO_X));
    GMG_TRY(smoothLevel(s, 0, cur0, alt0, b, false, upJacobi, false));
    // cur0 == x again after the second swap
    if (cur0 != x) GMG_TRY((launchVec<VO_COPY>(s, 0, x, cur0, nullptr, nullptr, 0, nullptr, KC_BLAS1, 16.0)));
    return GMG_OK;
}

// Replays `body` (a fixed sequence of kernel launches on the context's stream) from a cached CUDA graph.
template <typename Body>
static int runGraphed(gmg_solver *s, int kind, const void *p0, const void *p1, int flag, const Body &body)
{
    gmg_ctx *ctx = s->ctx;
    if (!s->useGraphs) return body();
    // profiling: a second variant of the graph with an external event-record node before and after every launch, so the
    // per-kernel times are those of the real replayed pipeline (warm L2, back-to-back launches), not of isolated launches
    const bool prof = ctx->profiling;
    auto key = std::make_tuple(kind + (prof ? (ctx->profileGroups ? 2000 : 1000) : 0), p0, p1, flag);
    auto it = s->graphs.find(key);
    if (it == s->graphs.end())
    {
	if (s->graphs.size() >= 16) dropGraphs(s);
	gmg_solver::GraphEntry e;
	const int64_t before = ctx->launches, commBefore = ctx->commOps;
	if (prof) flushProfile(ctx);
	GMG_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
	ctx->capturing = true;
	const int st = body();
	ctx->capturing = false;
	cudaError_t ce = cudaStreamEndCapture(ctx->stream, &e.graph);
	if (prof) { e.recs.swap(ctx->recs); ctx->recs.clear(); }
	if (st != GMG_OK) { if (e.graph) cudaGraphDestroy(e.graph); return st; }
	GMG_CUDA(ce);
	e.kernels = ctx->launches - before;
	e.comms = ctx->commOps - commBefore;
	ctx->launches = before;
	ctx->commOps = commBefore;
	GMG_CUDA(cudaGraphInstantiate(&e.exec, e.graph, 0));
	it = s->graphs.emplace(key, e).first;
    }
    GMG_CUDA(cudaGraphLaunch(it->second.exec, ctx->stream));
    ctx->launches += it->second.kernels;
    ctx->commOps += it->second.comms;
    if (prof)
    {
	GMG_CUDA(cudaStreamSynchronize(ctx->stream));
	accumulateRecs(ctx, it->second.recs, false);
    }
    return GMG_OK;
}

// L2 persistence for the level-0 band slab (buildBand): the stream's access-policy window marks accesses inside the slab as
// persisting, so the band metadata survives the full-grid passes that stream 100+ MB through L2 between two sweep groups.
// The attribute is a property of the stream (and is captured into the kernel nodes of the graphs), so it is (re)applied
// whenever another solver last used the context.  MEASURED (profiles/r02_ab_switches.md): the set-aside costs the full-grid
// passes more L2 than the band sweeps gain -- 256^3 solve 10.53 ms with the window, 10.25 ms without; 512^3 V-cycle 6.09 vs
// 5.4 ms -- so it is OFF unless GMG_L2_PERSIST=1 (or =<MB of set-aside>) asks for it.
static int applyBandWindow(gmg_solver *s)
{
    gmg_ctx *ctx = s->ctx;
    if (ctx->windowOwner == s) return GMG_OK;
    static const bool enabled = [] { const char *e = getenv("GMG_L2_PERSIST"); return e && e[0] != '0'; }();
    ctx->windowOwner = s;
    if (!enabled || ctx->persistBytes == 0) return GMG_OK;
    const Level &L = s->lv[0];
    cudaStreamAttrValue attr = {};
    if (L.bandSlab && L.nBand > 0)
    {
	const size_t bytes = std::min(L.bandSlabBytes, ctx->maxWindowBytes);
	attr.accessPolicyWindow.base_ptr = L.bandSlab;
	attr.accessPolicyWindow.num_bytes = bytes;
	attr.accessPolicyWindow.hitRatio = float(std::min(1.0, double(ctx->persistBytes) / double(std::max<size_t>(bytes, 1))));
	attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    GMG_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    return GMG_OK;
}

// doPrintStats (MG.h:24, the timing lines of MG.cpp:432-879): the V-cycle is launched kernel by kernel with a CUDA-event pair
// around every launch and the stages are printed under the reference's own headings -- device times, in seconds like
// UT_StopWatch's.  Consecutive band sweeps are one "Boundary smoother" line; the levels that run inside the fused coarse
// cycle (one kernel: no per-stage boundary exists) are one line.
static int vcyclePrintStats(gmg_solver *s, double *x, const double *b, bool useInitialGuess)
{
    gmg_ctx *ctx = s->ctx;
    flushProfile(ctx);
    const bool wasProfiling = ctx->profiling;
    ctx->profiling = true;
    const int st = vcycleLaunches(s, x, b, useInitialGuess);
    ctx->profiling = wasProfiling;
    cudaStreamSynchronize(ctx->stream);
    std::vector<ProfileRec> recs;
    recs.swap(ctx->recs);
    int heading = -100;
    bool up = false;
    auto ms = [](const ProfileRec &r) { float t = 0; cudaEventElapsedTime(&t, r.e0, r.e1); return double(t); };
    for (size_t i = 0; i < recs.size();)
    {
	const ProfileRec &r = recs[i];
	double t = ms(r);
	size_t j = i + 1;
	if (r.klass == KC_BAND || r.klass == KC_GS)
	    for (; j < recs.size() && recs[j].klass == r.klass && recs[j].level == r.level; ++j) t += ms(recs[j]);
	if (r.klass == KC_COARSE) up = true;
	const int lvl = r.klass == KC_RESTRICT ? r.level - 1 : r.level;  // a restriction is timed on the level it leaves (MG.cpp:536-552)
	const int key = (up && r.klass != KC_COARSE ? 1000 : 0) + lvl;
	if (r.klass != KC_COARSE && r.klass != KC_ZERO && r.klass != KC_BLAS1 && r.klass != KC_HALO && key != heading)
	{
	    heading = key;
	    if (lvl == 0) printf(up ? "    Fine Upstroke Smoother\n" : "    Fine Downstroke Smoother\n");
	    else printf(up ? "    Upstroke Smoother level: %d\n" : "    Downstroke Smoother level: %d\n", lvl);
	}
	switch (r.klass)
	{
	case KC_BAND: printf("      Boundary smoother time: %g\n", t * 1e-3); break;
	case KC_JACOBI: case KC_GS: printf("      Smoother time: %g\n", t * 1e-3); break;
	case KC_RESIDUAL: printf("      Compute residual time: %g\n", t * 1e-3); break;
	case KC_RESTRICT: printf("      Restriction time: %g\n", t * 1e-3); break;
	case KC_PROLONG: printf("      Prolongation time: %g\n", t * 1e-3); break;
	case KC_COARSE:
	    if (s->fusedFirst > 0) printf("    Levels %d to %d (down-stroke, direct solve, up-stroke in one kernel) time: %g\n", s->fusedFirst, s->levels - 1, t * 1e-3);
	    else printf("      Direct solve time: %g\n", t * 1e-3);
	    break;
	default: break;
	}
	i = j;
    }
    fflush(stdout);
    accumulateRecs(ctx, recs, true);
    return st;
}

static int vcycleDevice(gmg_solver *s, double *x, const double *b, bool useInitialGuess)
{
    GMG_TRY(applyBandWindow(s));
    if (s->opt.print_stats && !s->ctx->capturing)
    {
	if (s->opt.operators_only && s->levels > 1) return invalid("this solver handle was created with operators_only: no coarse factor, no V-cycle");
	return vcyclePrintStats(s, x, b, useInitialGuess);
    }
    if (s->opt.operators_only && s->levels > 1) return invalid("this solver handle was created with operators_only: no coarse factor, no V-cycle");
    return runGraphed(s, 0, x, b, useInitialGuess ? 1 : 0, [&]() { return vcycleLaunches(s, x, b, useInitialGuess); });
}

// ====================================================================================================
// PCG (CG.h:11-207)
// ====================================================================================================
// Device-side convergence test of the PCG loop (CG.h:159-161, :198): records sqrt(|r|^2/|b|^2), decides whether another
// iteration runs and tells the graph's WHILE node.  One thread.
struct PcgLoopState
{
    double threshold, bb;
    int iteration;    // the index CG.h:198 prints
    int maxIt;
    int histCount, histCap;
};
__global__ void k_pcg_check(PcgLoopState *st, double *hist, const Scalars *sc, cudaGraphConditionalHandle handle)
{
    pdlEnter();
    const double rr = sc->rr;
    if (st->histCount < st->histCap) hist[st->histCount++] = sqrt(rr / st->bb);
    unsigned more = 1u;
    if (rr < st->threshold) more = 0u;                 // CG.h:161
    else if (++st->iteration >= st->maxIt) more = 0u;  // CG.h:100
    cudaGraphSetConditional(handle, more);
}

// owned-plane reduction into a device scalar, summed over the ranks on a sharded level 0
template <int OP>
static int reduceOwned(gmg_solver *s, double *y, const double *a, size_t scalarOffset, double bytesPerCell, int level = 0)
{
    const Level &L = s->lv[level];
    GMG_TRY((launchVec<OP>(s, level, y, a, nullptr, nullptr, 0, scalarPtr(s, scalarOffset), KC_REDUCE, bytesPerCell, clipDepth(L, 0))));
    if (L.sharded) GMG_TRY(allreduceScalar(s, scalarPtr(s, scalarOffset), OP == VO_MAX ? NCCL_MAX : NCCL_SUM));
    return GMG_OK;
}

// the grid the diagonal preconditioner multiplies by (GFS.cpp:487-560), built on first use
static int ensureDiagInverse(gmg_solver *s)
{
    if (s->diagInv) return GMG_OK;
    gmg_ctx *ctx = s->ctx;
    Level &L = s->lv[0];
    GMG_TRY(allocGrid(&s->diagInv, L.g));
    const unsigned grid = unsigned(L.nChunksActive + divUp(L.nBoundary, BLOCK));
    if (grid == 0) return GMG_OK;
    GMG_LAUNCH(ctx, KC_SETUP, 0);
    k_diag_inverse<<<grid, BLOCK, 0, ctx->stream>>>(s->diagInv, L.labels, L.chunksActive, L.nChunksActive, L.g.chunksPerPlane, L.g.plane, L.g.n[2],
						    L.bandIdx, L.bcoef, L.nBoundary);
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}


// ====================================================================================================
// Mixed precision (the reference's own TODO, README.md:34-35): the multigrid preconditioner in fp32 inside
// the fp64 CG.  The levels that run as kernels use the float instantiations of the same stencil / band / transfer bodies on
// fp32 copies of their grids (half the bytes of every V-cycle pass); the fused coarse cycle keeps its fp64 shared-memory
// arithmetic and converts at its top level.  r crosses into fp32 once, z back into fp64 once.  Single GPU, Jacobi smoother.
// ====================================================================================================
static int ensureMixed(gmg_solver *s)
{
    if (s->lv[0].x32) return GMG_OK;
    if (s->ctx->world > 1) return invalid("mixed_precision is available on single-GPU contexts only");
    if (s->opt.use_gauss_seidel) return invalid("mixed_precision needs the damped-Jacobi smoother");
    if (s->fusedFirst > 0 && !s->compactArgs) return invalid("mixed_precision does not combine with the cluster coarse cycle (GMG_CLUSTER_CYCLE=1)");
    if (s->opt.boundary_iterations < 1) return invalid("mixed_precision needs at least one boundary smoother iteration");
    if (s->levels < 2) return invalid("mixed_precision needs at least two levels");
    // levels [0, nReg) run as fp32 kernels; level nReg is the top of the fused coarse cycle, or the direct solve's level
    const int nReg = s->fusedFirst > 0 ? s->fusedFirst : s->levels - 1;
    for (int l = 0; l <= nReg; ++l)
    {
	Level &L = s->lv[l];
	GMG_TRY(allocGrid32(&L.x32, L.g));
	GMG_TRY(allocGrid32(&L.b32, L.g));
	if (l < nReg)
	{
	    GMG_TRY(allocGrid32(&L.xAlt32, L.g));
	    GMG_TRY(allocGrid32(&L.r32, L.g));
	    const size_t n1 = size_t(std::max(L.nBand, 1));
	    GMG_CUDA(devMalloc(&L.bandV0f, sizeof(float) * n1));
	    GMG_CUDA(devMalloc(&L.bandV1f, sizeof(float) * n1));
	    GMG_CUDA(devMalloc(&L.bandBf, sizeof(float) * n1));
	}
    }
    return GMG_OK;
}

static int launchStencil32(gmg_solver *s, int level, int mode, const float *in, const float *b, float *out)
{
    s->ctx->curLevel = level;
    const Level &L = s->lv[level];
    StencilArgsT<float> a;
    a.labels = L.labels; a.flags = L.nbrMask; a.in = in; a.b = b; a.out = out;
    a.chunks = L.chunksInterior; a.nChunks = L.nChunksInterior; a.chunksPerPlane = L.g.chunksPerPlane;
    a.pitch = L.g.pitch; a.plane = L.g.plane; a.nz = L.g.n[2]; a.zlo = 0; a.zhi = L.g.n[2]; a.dotLo = 0; a.dotHi = L.g.n[2];
    a.nBoundary = L.nBoundary; a.bandIdx = L.bandIdx; a.bcoef = L.bcoef; a.wcode = L.wcode;
    a.partials = nullptr; a.ticket = nullptr; a.result = nullptr;
    const unsigned grid = unsigned(L.nChunksInterior + divUp(L.nBoundary, BLOCK));
    if (grid == 0) return GMG_OK;
    cudaStream_t st = s->ctx->stream;
    const double n = double(L.nActive);
    if (mode == SM_JACOBI)
    {
	GMG_LAUNCH(s->ctx, KC_JACOBI, n * 13.0);
	GMG_CUDA(launchK((k_stencil<SM_JACOBI, false, float>), grid, unsigned(BLOCK), size_t(0), st, a));
    }
    else if (mode == SM_JACOBI_ZERO)
    {
	GMG_LAUNCH(s->ctx, KC_JACOBI, n * 10.0);
	GMG_CUDA(launchK((k_stencil<SM_JACOBI_ZERO, false, float>), grid, unsigned(BLOCK), size_t(0), st, a));
    }
    else
    {
	GMG_LAUNCH(s->ctx, KC_RESIDUAL, n * 13.0);
	GMG_CUDA(launchK((k_stencil<SM_RESIDUAL, false, float>), grid, unsigned(BLOCK), size_t(0), st, a));
    }
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

static int launchBand32(gmg_solver *s, int level, float *x, const float *b, int sweeps, bool zeroGrid)
{
    s->ctx->curLevel = level;
    const Level &L = s->lv[level];
    if (L.nBand == 0 || sweeps <= 0) return GMG_OK;
    BandArgsT<float> a;
    a.x = x; a.b = b; a.bandIdx = L.bandIdx; a.bandRef = L.bandRef; a.bcoef = L.bcoef; a.wcode = L.wcode; a.bandB = L.bandBf;
    a.nBoundary = L.nBoundary; a.nBand = L.nBand; a.pitch = L.g.pitch; a.plane = L.g.plane;
    const unsigned grid = unsigned(divUp(L.nBand, BLOCK * BAND_PER_THREAD));
    cudaStream_t st = s->ctx->stream;
    const double bytes = double(L.nBand) * 17.0;
    const bool hw = L.hasWeights;
    float *cur = L.bandV0f, *nxt = L.bandV1f;
    a.vin = nullptr;
    a.vout = cur;
    {
	GMG_LAUNCH(s->ctx, KC_BAND, bytes);
	if (zeroGrid) GMG_CUDA(launchK((k_band<false, false, true, true, false, false, float>), grid, BLOCK, 0, st, a));
	else if (hw) GMG_CUDA(launchK((k_band<false, false, true, false, true, false, float>), grid, BLOCK, 0, st, a));
	else GMG_CUDA(launchK((k_band<false, false, true, false, false, false, float>), grid, BLOCK, 0, st, a));
    }
    if (sweeps == 1)
    {
	GMG_LAUNCH(s->ctx, KC_BAND, double(L.nBand) * 12.0);
	GMG_CUDA(launchK(k_band_scatter<float>, unsigned(divUp(L.nBand, BLOCK)), BLOCK, 0, st, x, L.bandIdx, static_cast<const float *>(cur), L.nBand));
    }
    for (int sw = 2; sw <= sweeps; ++sw)
    {
	a.vin = cur;
	a.vout = nxt;
	GMG_LAUNCH(s->ctx, KC_BAND, bytes);
	if (sw == sweeps)
	{
	    if (zeroGrid && hw) GMG_CUDA(launchK((k_band<true, true, false, false, true, true, float>), grid, BLOCK, 0, st, a));
	    else if (zeroGrid) GMG_CUDA(launchK((k_band<true, true, false, false, false, true, float>), grid, BLOCK, 0, st, a));
	    else if (hw) GMG_CUDA(launchK((k_band<true, true, false, false, true, false, float>), grid, BLOCK, 0, st, a));
	    else GMG_CUDA(launchK((k_band<true, true, false, false, false, false, float>), grid, BLOCK, 0, st, a));
	}
	else
	{
	    if (zeroGrid && hw) GMG_CUDA(launchK((k_band<true, false, false, false, true, true, float>), grid, BLOCK, 0, st, a));
	    else if (zeroGrid) GMG_CUDA(launchK((k_band<true, false, false, false, false, true, float>), grid, BLOCK, 0, st, a));
	    else if (hw) GMG_CUDA(launchK((k_band<true, false, false, false, true, false, float>), grid, BLOCK, 0, st, a));
	    else GMG_CUDA(launchK((k_band<true, false, false, false, false, false, float>), grid, BLOCK, 0, st, a));
	}
	std::swap(cur, nxt);
    }
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

static TransferArgsT<float> transferArgs32(gmg_solver *s, int fineLevel)
{
    const Level &F = s->lv[fineLevel], &C = s->lv[fineLevel + 1];
    TransferArgsT<float> a;
    a.fineLabels = F.labels; a.coarseLabels = C.labels;
    a.finePitch = F.g.pitch; a.coarsePitch = C.g.pitch; a.finePlane = F.g.plane; a.coarsePlane = C.g.plane;
    a.fineNz = F.g.n[2]; a.coarseNz = C.g.n[2]; a.coarseNy = C.g.n[1];
    for (int k = 0; k < 3; ++k) a.shift[k] = F.shift[k];
    a.fine = nullptr; a.coarse = nullptr; a.out = nullptr; a.chunks = nullptr; a.chunksPerPlane = 0;
    a.zlo = 0; a.zhi = 0;
    return a;
}

// one V-cycle in fp32: x32 = M^-1 b32 on level 0's fp32 grids (zero initial guess; the structure of vcycleLaunches)
static int vcycleLaunches32(gmg_solver *s)
{
    const int it = s->opt.boundary_iterations;
    const int nReg = s->fusedFirst > 0 ? s->fusedFirst : s->levels - 1;  // levels 0 .. nReg-1 run as kernels
    std::vector<float *> cur(std::max(nReg, 1)), alt(std::max(nReg, 1));
    auto smooth = [&](int level, bool zeroGrid) -> int {
	Level &L = s->lv[level];
	GMG_TRY(launchBand32(s, level, cur[level], L.b32, it, zeroGrid));
	GMG_TRY(launchStencil32(s, level, zeroGrid ? SM_JACOBI_ZERO : SM_JACOBI, cur[level], L.b32, alt[level]));
	std::swap(cur[level], alt[level]);
	GMG_TRY(launchBand32(s, level, cur[level], L.b32, it, false));
	return GMG_OK;
    };
    for (int level = 0; level < nReg; ++level)
    {
	Level &L = s->lv[level], &C = s->lv[level + 1];
	cur[level] = L.x32;
	alt[level] = L.xAlt32;
	GMG_TRY(smooth(level, true));
	GMG_TRY(launchStencil32(s, level, SM_RESIDUAL, cur[level], L.b32, L.r32));
	if (C.nChunksActive > 0)
	{
	    s->ctx->curLevel = level + 1;
	    TransferArgsT<float> a = transferArgs32(s, level);
	    a.fine = L.r32; a.out = C.b32; a.chunks = C.chunksActive; a.chunksPerPlane = C.g.chunksPerPlane; a.zlo = 0; a.zhi = C.g.n[2];
	    GMG_LAUNCH(s->ctx, KC_RESTRICT, double(L.nActive) * 4.0 + double(C.nActive) * 5.0);
	    GMG_CUDA(launchK(k_restrict<float>, unsigned(C.nChunksActive * RESTRICT_SPLIT), unsigned(BLOCK), size_t(0), s->ctx->stream, a));
	}
    }
    if (s->fusedFirst <= 0)
    {
	// no fused cycle: the direct solve of the coarsest level in fp64, through that level's fp64 grids
	Level &C = s->lv[s->levels - 1];
	s->ctx->curLevel = s->levels - 1;
	if (C.nChunksActive > 0)
	{
	    GMG_LAUNCH(s->ctx, KC_BLAS1, double(C.nActive) * 12.0);
	    GMG_CUDA(launchK((k_convert<double, float>), unsigned(C.nChunksActive), unsigned(BLOCK), size_t(0), s->ctx->stream, C.b, static_cast<const float *>(C.b32),
			     C.chunksActive, C.g.chunksPerPlane, C.g.plane, C.g.n[2]));
	}
	GMG_TRY(launchCoarse(s, C.x, C.b));
	if (C.nChunksActive > 0)
	{
	    GMG_LAUNCH(s->ctx, KC_BLAS1, double(C.nActive) * 12.0);
	    GMG_CUDA(launchK((k_convert<float, double>), unsigned(C.nChunksActive), unsigned(BLOCK), size_t(0), s->ctx->stream, C.x32, static_cast<const double *>(C.x),
			     C.chunksActive, C.g.chunksPerPlane, C.g.plane, C.g.n[2]));
	}
    }
    else
    {
	// the fused coarse cycle, fp64 inside, reading / writing the fp32 grids of its top level
	s->ctx->curLevel = s->fusedFirst;
	CompactArgs c = *static_cast<const CompactArgs *>(s->compactArgs);
	c.bTop32 = s->lv[s->fusedFirst].b32;
	c.xTop32 = s->lv[s->fusedFirst].x32;
	double bytes = 0;
	for (int l = s->fusedFirst; l < s->levels - 1; ++l) bytes += double(s->lv[l].nActive) * 126.0;
	GMG_LAUNCH(s->ctx, KC_COARSE, bytes);
	GMG_CUDA(launchK(k_compact_cycle, unsigned(1), unsigned(CYCLE_THREADS), size_t(s->compactSmem), s->ctx->stream, c));
    }
    for (int level = nReg - 1; level >= 0; --level)
    {
	Level &L = s->lv[level], &C = s->lv[level + 1];
	if (L.nChunksActive > 0)
	{
	    s->ctx->curLevel = level;
	    TransferArgsT<float> a = transferArgs32(s, level);
	    // the level below hands over its result in its x32 grid (two Jacobi swaps, or the fused cycle)
	    a.coarse = C.x32; a.out = cur[level]; a.chunks = L.chunksActive; a.chunksPerPlane = L.g.chunksPerPlane; a.zlo = 0; a.zhi = L.g.n[2];
	    GMG_LAUNCH(s->ctx, KC_PROLONG, double(L.nActive) * 9.0 + double(C.nActive) * 4.0);
	    GMG_CUDA(launchK(k_prolong<float>, unsigned(L.nChunksActive), unsigned(BLOCK), size_t(0), s->ctx->stream, a));
	}
	GMG_TRY(smooth(level, false));
    }
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;  // two swaps per level: every level's result is back in its x32
}

// z = M^-1 r with the fp32 V-cycle: r -> fp32, V-cycle, -> fp64
static int mixedPreconditioner(gmg_solver *s, double *z, const double *r)
{
    GMG_TRY(ensureMixed(s));
    gmg_ctx *ctx = s->ctx;
    Level &L = s->lv[0];
    if (L.nChunksActive == 0) return GMG_OK;
    ctx->curLevel = 0;
    {
	GMG_LAUNCH(ctx, KC_BLAS1, double(L.nActive) * 12.0);
	GMG_CUDA(launchK((k_convert<float, double>), unsigned(L.nChunksActive), unsigned(BLOCK), size_t(0), ctx->stream, L.b32, r, L.chunksActive, L.g.chunksPerPlane,
			 L.g.plane, L.g.n[2]));
    }
    GMG_TRY(vcycleLaunches32(s));
    {
	ctx->curLevel = 0;
	GMG_LAUNCH(ctx, KC_BLAS1, double(L.nActive) * 12.0);
	GMG_CUDA(launchK((k_convert<double, float>), unsigned(L.nChunksActive), unsigned(BLOCK), size_t(0), ctx->stream, z, static_cast<const float *>(L.x32), L.chunksActive,
			 L.g.chunksPerPlane, L.g.plane, L.g.n[2]));
    }
    GMG_CUDA(cudaGetLastError());
    return GMG_OK;
}

// z = M^-1 r for the three preconditioners of the ABI: 0 none (plain CG), 1 multigrid V-cycle (GFS.cpp:468-472),
// 2 diagonal (GFS.cpp:562-603)
static int applyPreconditioner(gmg_solver *s, int precond, double *z, const double *r, bool direct)
{
    const ZRange own = clipDepth(s->lv[0], 0);
    if (precond == 1 && s->opt.mixed_precision)
    {
	if (direct || !s->useGraphs) return mixedPreconditioner(s, z, r);
	GMG_TRY(ensureMixed(s));  // allocations stay outside the capture
	return runGraphed(s, 4, z, r, 0, [&]() { return mixedPreconditioner(s, z, r); });
    }
    if (precond == 1) return direct ? vcycleLaunches(s, z, r, false) : vcycleDevice(s, z, r, false);
    if (precond == 2) return launchVec<VO_MUL>(s, 0, z, r, s->diagInv, nullptr, 0, nullptr, KC_BLAS1, 24.0, own);
    return launchVec<VO_COPY>(s, 0, z, r, nullptr, nullptr, 0, nullptr, KC_BLAS1, 16.0, own);
}

static bool useDeviceLoop()
{
    static const bool v = [] { const char *e = getenv("GMG_DEVICE_LOOP"); return !(e && e[0] == '0'); }();
    return v;
}

// Sharded level 0: x0 and b come in valid over the whole stored slab.  Each iteration the search
// direction p is refreshed HALO_P = 9 planes deep, so t = A p is valid 8 deep and the update r -= alpha t keeps r valid
// 8 deep redundantly -- exactly what the next V-cycle needs as its right-hand side -- without an exchange of its own.
static int pcgDevice(gmg_solver *s, double *x, const double *b, double tol, int maxIt, int precond, int *iterations, double *hist, int histCap,
		     int *histCount)
{
    gmg_ctx *ctx = s->ctx;
    const Level &L0 = s->lv[0];
    GMG_TRY(applyBandWindow(s));
    if (precond < 0 || precond > 2) return invalid("gmg_pcg: preconditioner must be 0 (none), 1 (multigrid V-cycle) or 2 (diagonal)");
    if (!s->pcgR)
    {
	GMG_TRY(allocGrid(&s->pcgR, s->lv[0].g));
	GMG_TRY(allocGrid(&s->pcgP, s->lv[0].g));
	GMG_TRY(allocGrid(&s->pcgZ, s->lv[0].g));
	GMG_TRY(allocGrid(&s->pcgT, s->lv[0].g));
    }
    double *r = s->pcgR, *p = s->pcgP, *z = s->pcgZ, *t = s->pcgT;
    const ZRange own = clipDepth(L0, 0), deep = clipDepth(L0, HALO_X);
    if (histCount) *histCount = 0;
    if (iterations) *iterations = -1;
    double bb = 0, rr = 0;
    GMG_TRY((reduceOwned<VO_NORM2>(s, const_cast<double *>(b), nullptr, offsetof(Scalars, bb), 8.0)));
    GMG_TRY(readScalar(s, offsetof(Scalars, bb), &bb));
    if (bb == 0) return GMG_OK; // "RHS is zero. Nothing to solve" (CG.h:35-40)
    // r = b - A x (CG.h:50-51)
    GMG_TRY(launchStencil(s, 0, SM_RESIDUAL, x, b, r, nullptr, deep));
    GMG_TRY((reduceOwned<VO_NORM2>(s, r, nullptr, offsetof(Scalars, rr), 8.0)));
    GMG_TRY(readScalar(s, offsetof(Scalars, rr), &rr));
    const double threshold = tol * tol * bb;
    if (rr < threshold) return GMG_OK; // CG.h:60-64
    if (maxIt <= 0)
    {
	// CG.h:100: the loop body never runs -- x is left untouched and the index printed at CG.h:198 is 0 (the reference
	// still applies the preconditioner once for a search direction nobody uses; that has no visible effect)
	if (iterations) *iterations = 0;
	return GMG_OK;
    }
    // p = M^-1 r ; rho = p.r (CG.h:66-87)
    // The first solve of a solver launches its one-off pieces (first preconditioner application) directly: a solver is
    // typically built, used for ONE solve and destroyed every simulation frame, and capturing + instantiating a graph that
    // is replayed once costs more than it saves.  From the second solve on they are cached graphs like the rest.
    if (precond == 1 && s->opt.operators_only && s->levels > 1) return invalid("this solver handle was created with operators_only: no coarse factor, no V-cycle");
    if (precond == 2) GMG_TRY(ensureDiagInverse(s));
    if (precond == 1 && s->opt.mixed_precision) GMG_TRY(ensureMixed(s));
    // (single GPU only: on a sharded context every exchange stays inside the graphs it was validated in)
    const bool firstSolve = (s->pcgSolves++ == 0) && !ctx->profiling && ctx->world == 1;
    GMG_TRY(applyPreconditioner(s, precond, p, r, firstSolve));
    GMG_TRY((reduceOwned<VO_DOT>(s, p, r, offsetof(Scalars, rhoNew), 16.0)));
    // Reference loop (CG.h:100-195): [t = A p, alpha, x += alpha p, r -= alpha t, |r|^2, test] then
    // [z = M^-1 r, beta, p = z + beta p].  Re-bracketed here as: first [apply, update]; then per iteration
    // {preconditioner, z.r, direction, apply (+p.Ap), update (+|r|^2, retires rho)} and the test.
    auto applyUpdate = [&]() -> int {
	GMG_TRY(haloExchange(s, 0, p, HALO_P));
	// sharded: t = A p and the update run over the deep halo range (r stays valid 8 planes deep), their fused reductions
	// only sum the rank's owned planes and are then summed over the ranks
	GMG_TRY(launchStencil(s, 0, SM_APPLY, p, nullptr, t, scalarPtr(s, offsetof(Scalars, pAp)), deep));
	if (L0.sharded) GMG_TRY(allreduceScalar(s, scalarPtr(s, offsetof(Scalars, pAp))));
	GMG_TRY((launchVec<VO_CG_UPDATE>(s, 0, x, p, t, r, 0, scalarPtr(s, offsetof(Scalars, rr)), KC_BLAS1, 48.0, deep)));
	if (L0.sharded) GMG_TRY(allreduceScalar(s, scalarPtr(s, offsetof(Scalars, rr))));
	return GMG_OK;
    };
    auto iterationBody = [&]() -> int {
	GMG_TRY(applyPreconditioner(s, precond, z, r, true));
	GMG_TRY((reduceOwned<VO_DOT>(s, z, r, offsetof(Scalars, rhoNew), 16.0)));
	GMG_TRY((launchVec<VO_CG_DIRECTION>(s, 0, p, z, nullptr, nullptr, 0, nullptr, KC_BLAS1, 24.0, own)));
	return applyUpdate();
    };
    int iteration = 0;
    if (useDeviceLoop() && s->useGraphs && !ctx->profiling && ctx->world == 1 && !ctx->deviceLoopBroken)
    {
	// The whole loop as ONE graph launch: [apply, update, test] then a WHILE node around {iteration, test}; the residual
	// history stays on the device until the loop ends, so the host synchronises once per solve instead of once per
	// iteration.  The loop parameters live in device memory, so the cached graph serves every (tol, maxIt).
	constexpr int HIST_MAX = 4096;
	if (!s->pcgLoop)
	{
	    char *blk = nullptr;
	    GMG_CUDA(devMalloc(&blk, sizeof(PcgLoopState) + sizeof(double) * HIST_MAX));
	    s->pcgLoop = blk;
	    GMG_CUDA(cudaMallocHost(&s->pcgLoopHost, sizeof(PcgLoopState) + sizeof(double) * HIST_MAX));
	}
	PcgLoopState *dst = static_cast<PcgLoopState *>(s->pcgLoop);
	double *dhist = reinterpret_cast<double *>(dst + 1);
	PcgLoopState *hst = static_cast<PcgLoopState *>(s->pcgLoopHost);
	auto key = std::make_tuple(3, static_cast<const void *>(x), static_cast<const void *>(b), precond);
	auto it = s->graphs.find(key);
	if (it == s->graphs.end())
	{
	    if (s->graphs.size() >= 16) dropGraphs(s);
	    gmg_solver::GraphEntry e;
	    const int64_t before = ctx->launches;
	    Scalars *sc = reinterpret_cast<Scalars *>(ctx->scalars);
	    cudaGraphConditionalHandle handle = 0;
	    auto check = [&]() -> int {
		ctx->curLevel = 0;
		GMG_LAUNCH(ctx, KC_BLAS1, 0);
		GMG_CUDA(launchK(k_pcg_check, unsigned(1), unsigned(1), size_t(0), ctx->stream, dst, dhist, static_cast<const Scalars *>(sc), handle));
		return GMG_OK;
	    };
	    // returns GMG_OK with e.exec set, or an error with the stream out of capture mode and nothing leaked
	    auto build = [&]() -> int {
		GMG_CUDA(cudaGraphCreate(&e.graph, 0));
		GMG_CUDA(cudaGraphConditionalHandleCreate(&handle, e.graph, 0, cudaGraphCondAssignDefault));
		// head: [apply, update, test]
		GMG_CUDA(cudaStreamBeginCaptureToGraph(ctx->stream, e.graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		ctx->capturing = true;
		int st = applyUpdate();
		if (st == GMG_OK) st = check();
		ctx->capturing = false;
		cudaStreamCaptureStatus cs;
		const cudaGraphNode_t *deps = nullptr;
		size_t nDeps = 0;
		const cudaError_t ce = cudaStreamGetCaptureInfo(ctx->stream, &cs, nullptr, nullptr, &deps, &nDeps);
		std::vector<cudaGraphNode_t> tail;
		if (ce == cudaSuccess) tail.assign(deps, deps + nDeps);
		cudaGraph_t g1 = nullptr;
		const cudaError_t ce2 = cudaStreamEndCapture(ctx->stream, &g1);
		e.kernels = ctx->launches - before;
		if (st != GMG_OK) return st;
		GMG_CUDA(ce);
		GMG_CUDA(ce2);
		// WHILE node around {iteration, test}
		cudaGraphNodeParams np = {};
		np.type = cudaGraphNodeTypeConditional;
		np.conditional.handle = handle;
		np.conditional.type = cudaGraphCondTypeWhile;
		np.conditional.size = 1;
		cudaGraphNode_t whileNode;
		GMG_CUDA(cudaGraphAddNode(&whileNode, e.graph, tail.data(), tail.size(), &np));
		cudaGraph_t body = np.conditional.phGraph_out[0];
		const int64_t beforeBody = ctx->launches;
		GMG_CUDA(cudaStreamBeginCaptureToGraph(ctx->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
		ctx->capturing = true;
		st = iterationBody();
		if (st == GMG_OK) st = check();
		ctx->capturing = false;
		cudaGraph_t g2 = nullptr;
		const cudaError_t ce3 = cudaStreamEndCapture(ctx->stream, &g2);
		e.comms = ctx->launches - beforeBody;  // (re-used field: kernels per loop iteration)
		if (st != GMG_OK) return st;
		GMG_CUDA(ce3);
		GMG_CUDA(cudaGraphInstantiate(&e.exec, e.graph, 0));
		return GMG_OK;
	    };
	    const int st = build();
	    ctx->launches = before;
	    if (st != GMG_OK)
	    {
		// a driver that refuses the conditional graph: remember it and run the host loop below (nothing has executed yet)
		if (e.graph) cudaGraphDestroy(e.graph);
		cudaGetLastError();
		ctx->deviceLoopBroken = true;
		if (getenv("GMG_TRACE")) fprintf(stderr, "[gmg] device-side PCG loop unavailable (%s): host loop\n", gmg_last_error());
	    }
	    else it = s->graphs.emplace(key, e).first;
	}
	if (it != s->graphs.end())
	{
	    hst->threshold = threshold; hst->bb = bb; hst->iteration = 0; hst->maxIt = maxIt; hst->histCount = 0;
	    hst->histCap = std::min(HIST_MAX, std::max(maxIt + 1, 1));
	    const int cap = hst->histCap;
	    GMG_CUDA(cudaMemcpyAsync(dst, hst, sizeof(PcgLoopState), cudaMemcpyHostToDevice, ctx->stream));
	    GMG_CUDA(cudaGraphLaunch(it->second.exec, ctx->stream));
	    GMG_CUDA(cudaMemcpyAsync(hst, dst, sizeof(PcgLoopState) + sizeof(double) * cap, cudaMemcpyDeviceToHost, ctx->stream));
	    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
	    iteration = hst->iteration;
	    ctx->launches += it->second.kernels + int64_t(std::max(0, hst->histCount - 1)) * it->second.comms;
	    const double *hh = reinterpret_cast<const double *>(hst + 1);
	    if (hist && histCount)
		for (int k = 0; k < hst->histCount && *histCount < histCap; ++k) hist[(*histCount)++] = hh[k];
	    if (iterations) *iterations = iteration;
	    return GMG_OK;
	}
    }
    if (firstSolve) GMG_TRY(applyUpdate());
    else GMG_TRY(runGraphed(s, 1, x, nullptr, 0, applyUpdate));
    for (;;)
    {
	GMG_TRY(readScalar(s, offsetof(Scalars, rr), &rr));
	if (hist && histCount && *histCount < histCap) hist[(*histCount)++] = std::sqrt(rr / bb);
	if (rr < threshold) break;
	if (++iteration >= maxIt) break; // CG.h:198 then prints maxIterations
	GMG_TRY(runGraphed(s, 2, x, nullptr, precond, iterationBody));
    }
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (iterations) *iterations = iteration;
    return p2pCheckError(ctx);
}

// ====================================================================================================
// grids and the public operator entry points
// ====================================================================================================
extern "C" int gmg_grid_create(gmg_solver *s, int level, gmg_grid **out)
{
    if (!s || !out || level < 0 || level >= s->levels) return invalid("gmg_grid_create: bad argument");
    GMG_CUDA(enterCtx(s->ctx));
    gmg_grid *g = new gmg_grid;
    g->solver = s;
    g->level = level;
    int st = allocGrid(&g->d, s->lv[level].g);
    if (st != GMG_OK) { delete g; return st; }
    g->plane = s->lv[level].g.plane;
    *out = g;
    return GMG_OK;
}
extern "C" int gmg_grid_destroy(gmg_grid *g)
{
    if (!g) return GMG_OK;
    enterCtx(g->solver->ctx);  // the free is ordered on THIS context's stream, behind the work that still uses the grid
    if (g->d) devFree(g->d - g->plane);
    delete g;
    return GMG_OK;
}
extern "C" int gmg_grid_upload(gmg_grid *g, const double *host)
{
    if (!g || !host) return invalid("null argument");
    GMG_CUDA(enterCtx(g->solver->ctx));
    const Geom &ge = g->solver->lv[g->level].g;
    return uploadValues(g->solver->ctx, g->d, host, ge.res, ge, g->solver->lv[g->level].labels, g->level == 0 ? g->solver->hostBounds : nullptr);
}
// the rank's owned planes as a box of their own (what a sharded context hands back to the host)
static Geom ownedGeom(const Level &L)
{
    Geom go = L.g;
    go.org[2] = L.g.org[2] + L.ownLo;
    go.n[2] = L.ownHi - L.ownLo;
    go.total = go.plane * go.n[2];
    go.zBlocks = int(divUp(go.n[2], CHUNK_Z));
    return go;
}
extern "C" int gmg_grid_download(gmg_grid *g, double *host)
{
    if (!g || !host) return invalid("null argument");
    GMG_CUDA(enterCtx(g->solver->ctx));
    const Level &L = g->solver->lv[g->level];
    const Geom go = ownedGeom(L);
    GMG_TRY(downloadValues(g->solver->ctx, host, g->d + int64_t(L.ownLo) * L.g.plane, go.res, go, true));
    return p2pCheckError(g->solver->ctx);  // sticky: a timed-out exchange anywhere before this download
}
extern "C" int gmg_grid_zero(gmg_grid *g)
{
    if (!g) return invalid("null argument");
    GMG_CUDA(enterCtx(g->solver->ctx));
    GMG_CUDA(cudaMemsetAsync(g->d, 0, sizeof(double) * g->solver->lv[g->level].g.total, g->solver->ctx->stream));
    return GMG_OK;
}
extern "C" int gmg_grid_copy(gmg_grid *dst, const gmg_grid *src)
{
    if (!dst || !src || dst->level != src->level || dst->solver != src->solver) return invalid("gmg_grid_copy: grids differ in level");
    GMG_CUDA(enterCtx(dst->solver->ctx));
    GMG_CUDA(cudaMemcpyAsync(dst->d, src->d, sizeof(double) * dst->solver->lv[dst->level].g.total, cudaMemcpyDeviceToDevice, dst->solver->ctx->stream));
    return GMG_OK;
}

#define CHECK_GRIDS2(s, a, b)                                                                  \
    if (!(s) || !(a) || !(b) || (a)->solver != (s) || (b)->solver != (s) || (a)->level != (b)->level) \
    return invalid("grid arguments must belong to this solver and share a level")

// On a sharded level the single-operator entry points first refresh the halo planes their stencil reads, then compute
// the rank's owned planes; reductions sum the owned planes of every rank.
static int checkSweeps(const Level &L, int sweeps)
{
    if (L.sharded && sweeps > HALO_STORE) return invalid("at most 8 boundary sweeps per call on a sharded level (stored halo depth)");
    return GMG_OK;
}

extern "C" int gmg_jacobi(gmg_solver *s, gmg_grid *x, const gmg_grid *b)
{
    CHECK_GRIDS2(s, x, b);
    GMG_CUDA(enterCtx(s->ctx));
    Level &L = s->lv[x->level];
    GMG_TRY(haloExchange(s, x->level, x->d, 1));
    GMG_TRY(launchStencil(s, x->level, SM_JACOBI, x->d, b->d, L.xAlt, nullptr, clipDepth(L, 0)));
    // copy back rather than swap pointers: cached V-cycle graphs hold the level's xAlt address
    return launchVec<VO_COPY>(s, x->level, x->d, L.xAlt, nullptr, nullptr, 0, nullptr, KC_BLAS1, 16.0, clipDepth(L, 0));
}
extern "C" int gmg_gauss_seidel(gmg_solver *s, gmg_grid *x, const gmg_grid *b, int oddTiles, int forward)
{
    CHECK_GRIDS2(s, x, b);
    GMG_CUDA(enterCtx(s->ctx));
    return launchGaussSeidel(s, x->level, x->d, b->d, oddTiles != 0, forward != 0);
}
extern "C" int gmg_boundary_jacobi(gmg_solver *s, gmg_grid *x, const gmg_grid *b, int sweeps)
{
    CHECK_GRIDS2(s, x, b);
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY(checkSweeps(s->lv[x->level], sweeps));
    GMG_TRY(haloExchange(s, x->level, x->d, sweeps));
    GMG_TRY(haloExchange(s, x->level, b->d, std::max(0, sweeps - 1)));
    return launchBand(s, x->level, x->d, b->d, sweeps, false);
}
extern "C" int gmg_apply(gmg_solver *s, gmg_grid *dst, const gmg_grid *src)
{
    CHECK_GRIDS2(s, dst, src);
    if (dst == src) return invalid("gmg_apply: dst must differ from src");
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY(haloExchange(s, src->level, src->d, 1));
    return launchStencil(s, dst->level, SM_APPLY, src->d, nullptr, dst->d, nullptr, clipDepth(s->lv[dst->level], 0));
}
extern "C" int gmg_residual(gmg_solver *s, gmg_grid *r, const gmg_grid *x, const gmg_grid *b)
{
    CHECK_GRIDS2(s, r, x);
    CHECK_GRIDS2(s, r, b);
    if (r == x) return invalid("gmg_residual: r must differ from x");
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY(haloExchange(s, x->level, x->d, 1));
    return launchStencil(s, r->level, SM_RESIDUAL, x->d, b->d, r->d, nullptr, clipDepth(s->lv[r->level], 0));
}
extern "C" int gmg_restrict(gmg_solver *s, gmg_grid *coarse, const gmg_grid *fine)
{
    if (!s || !coarse || !fine || coarse->solver != s || fine->solver != s || coarse->level != fine->level + 1)
	return invalid("gmg_restrict: coarse must be one level above fine");
    GMG_CUDA(enterCtx(s->ctx));
    const Level &F = s->lv[fine->level], &C = s->lv[coarse->level];
    GMG_TRY(haloExchange(s, fine->level, fine->d, 1));
    ZRange zr = {0, C.g.n[2]};
    if (C.sharded) zr = {C.ownLo, C.ownHi};
    else if (F.sharded) zr = {s->gatherLo[s->ctx->rank], s->gatherHi[s->ctx->rank]};
    GMG_TRY(launchRestrict(s, fine->level, coarse->d, fine->d, zr));
    if (F.sharded && !C.sharded) GMG_TRY(gatherReplicated(s, coarse->d));
    return GMG_OK;
}
extern "C" int gmg_prolong_add(gmg_solver *s, gmg_grid *fine, const gmg_grid *coarse)
{
    if (!s || !coarse || !fine || coarse->solver != s || fine->solver != s || coarse->level != fine->level + 1)
	return invalid("gmg_prolong_add: coarse must be one level above fine");
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY(haloExchange(s, coarse->level, coarse->d, 1));
    return launchProlong(s, fine->level, fine->d, coarse->d, clipDepth(s->lv[fine->level], 0));
}
extern "C" int gmg_dot(gmg_solver *s, const gmg_grid *a, const gmg_grid *b, double *out)
{
    CHECK_GRIDS2(s, a, b);
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY((reduceOwned<VO_DOT>(s, a->d, b->d, offsetof(Scalars, tmp), 16.0, a->level)));
    return readScalar(s, offsetof(Scalars, tmp), out);
}
extern "C" int gmg_norm2(gmg_solver *s, const gmg_grid *a, double *out)
{
    CHECK_GRIDS2(s, a, a);
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY((reduceOwned<VO_NORM2>(s, a->d, nullptr, offsetof(Scalars, tmp), 8.0, a->level)));
    return readScalar(s, offsetof(Scalars, tmp), out);
}
extern "C" int gmg_inf_norm(gmg_solver *s, const gmg_grid *a, double *out)
{
    CHECK_GRIDS2(s, a, a);
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY((reduceOwned<VO_MAX>(s, a->d, nullptr, offsetof(Scalars, tmp), 8.0, a->level)));
    return readScalar(s, offsetof(Scalars, tmp), out);
}
extern "C" int gmg_axpy(gmg_solver *s, gmg_grid *dst, const gmg_grid *src, double scale)
{
    CHECK_GRIDS2(s, dst, src);
    GMG_CUDA(enterCtx(s->ctx));
    return launchVec<VO_AXPY>(s, dst->level, dst->d, src->d, nullptr, nullptr, scale, nullptr, KC_BLAS1, 24.0);
}
extern "C" int gmg_add_scaled(gmg_solver *s, gmg_grid *dst, const gmg_grid *a, const gmg_grid *v, double scale)
{
    CHECK_GRIDS2(s, dst, a);
    CHECK_GRIDS2(s, dst, v);
    GMG_CUDA(enterCtx(s->ctx));
    return launchVec<VO_ADD_SCALED>(s, dst->level, dst->d, a->d, v->d, nullptr, scale, nullptr, KC_BLAS1, 24.0);
}
extern "C" int gmg_scale(gmg_solver *s, gmg_grid *v, double scale)
{
    CHECK_GRIDS2(s, v, v);
    GMG_CUDA(enterCtx(s->ctx));
    return launchVec<VO_SCALE>(s, v->level, v->d, nullptr, nullptr, nullptr, scale, nullptr, KC_BLAS1, 16.0);
}

// device-resident grids may have been produced by owned-plane operators: refresh the deep halo the solve starts from
static int refreshForSolve(gmg_solver *s, double *x, const double *b, bool xMatters)
{
    if (!s->lv[0].sharded) return GMG_OK;
    GMG_TRY(haloExchange(s, 0, const_cast<double *>(b), HALO_STORE0));
    if (xMatters) GMG_TRY(haloExchange(s, 0, x, HALO_STORE0));
    return GMG_OK;
}

extern "C" int gmg_vcycle_device(gmg_solver *s, gmg_grid *x, const gmg_grid *b, int useInitialGuess)
{
    CHECK_GRIDS2(s, x, b);
    if (x->level != 0) return invalid("gmg_vcycle_device: grids must be level 0");
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY(refreshForSolve(s, x->d, b->d, useInitialGuess != 0));
    return vcycleDevice(s, x->d, b->d, useInitialGuess != 0);
}
extern "C" int gmg_pcg_device(gmg_solver *s, gmg_grid *x, const gmg_grid *b, double tol, int maxIt, int preconditioner, int *iterations,
			      double *relResHistory, int histCap, int *histCount)
{
    CHECK_GRIDS2(s, x, b);
    if (x->level != 0) return invalid("gmg_pcg_device: grids must be level 0");
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY(refreshForSolve(s, x->d, b->d, true));
    return pcgDevice(s, x->d, b->d, tol, maxIt, preconditioner, iterations, relResHistory, histCap, histCount);
}

// communication micro-benchmark: `reps` back-to-back operations of one kind on a sharded solver (no compute in between,
// so no load skew): kind 0 = halo exchange of `depth` planes at `level`, 1 = gather of the first replicated level,
// 2 = scalar all-reduce.  Collective.  msPerOp: device time per operation on this rank.
extern "C" int gmg_comm_benchmark(gmg_solver *s, int kind, int level, int depth, int reps, double *msPerOp)
{
    if (!s || !msPerOp || reps < 1) return invalid("gmg_comm_benchmark: bad argument");
    if (s->shardLevels == 0) return invalid("gmg_comm_benchmark: the solver is not sharded");
    if (kind == 0 && (level < 0 || level >= s->shardLevels || depth < 1 || depth > HALO_STORE)) return invalid("gmg_comm_benchmark: bad level / depth");
    gmg_ctx *ctx = s->ctx;
    GMG_CUDA(enterCtx(ctx));
    auto once = [&]() -> int {
	if (kind == 0) return haloExchange(s, level, s->lv[level].r, depth);
	if (kind == 1) return gatherReplicated(s, s->lv[s->shardLevels].b);
	return allreduceScalar(s, scalarPtr(s, offsetof(Scalars, tmp)));
    };
    for (int i = 0; i < 3; ++i) GMG_TRY(once());
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    GMG_CUDA(cudaEventRecord(ctx->t0, ctx->stream));
    for (int i = 0; i < reps; ++i) GMG_TRY(once());
    GMG_CUDA(cudaEventRecord(ctx->t1, ctx->stream));
    GMG_CUDA(cudaEventSynchronize(ctx->t1));
    float ms = 0;
    GMG_CUDA(cudaEventElapsedTime(&ms, ctx->t0, ctx->t1));
    *msPerOp = double(ms) / reps;
    return p2pCheckError(ctx);
}

static int ensureHostIO(gmg_solver *s)
{
    if (s->pcgX) return GMG_OK;
    GMG_TRY(allocGrid(&s->pcgX, s->lv[0].g));
    GMG_TRY(allocGrid(&s->pcgB, s->lv[0].g));
    return GMG_OK;
}

extern "C" int gmg_vcycle(gmg_solver *s, double *x, const double *b, int useInitialGuess)
{
    if (!s || !x || !b) return invalid("gmg_vcycle: null argument");
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY(ensureHostIO(s));
    const Geom &g = s->lv[0].g;
    if (useInitialGuess) GMG_TRY(uploadValues(s->ctx, s->pcgX, x, g.res, g, s->lv[0].labels, s->hostBounds, &s->ioGroups));
    GMG_TRY(uploadValues(s->ctx, s->pcgB, b, g.res, g, s->lv[0].labels, s->hostBounds, &s->ioGroups));
    GMG_TRY(vcycleDevice(s, s->pcgX, s->pcgB, useInitialGuess != 0));
    // cells outside the stored box are non-active: the reference leaves them untouched (0 after constant(0));
    // a sharded context writes the rank's owned planes only
    const Geom go = ownedGeom(s->lv[0]);
    return downloadValues(s->ctx, x, s->pcgX + int64_t(s->lv[0].ownLo) * g.plane, g.res, go, !useInitialGuess, s->hostBounds, &s->ioGroups, s->lv[0].ownLo);
}

static int pcgHost(gmg_solver *s, double *x, const double *b, double tol, int maxIt, int preconditioner, int *iterations, double *relResHistory, int histCap,
		   int *histCount, bool fromZero)
{
    if (!s || !x || !b) return invalid("gmg_pcg: null argument");
    GMG_CUDA(enterCtx(s->ctx));
    GMG_TRY(ensureHostIO(s));
    const Geom &g = s->lv[0].g;
    // fromZero: the caller declares x == 0 on entry (the node's solutionGrid.constant(0) without a warm start, GFS.cpp:392-398): nothing of
    // it crosses PCIe, the device grid is cleared instead; the arithmetic is the same (r = b - A 0)
    if (fromZero) GMG_CUDA(cudaMemsetAsync(s->pcgX, 0, sizeof(double) * g.total, s->ctx->stream));
    else GMG_TRY(uploadValues(s->ctx, s->pcgX, x, g.res, g, s->lv[0].labels, s->hostBounds, &s->ioGroups));
    GMG_TRY(uploadValues(s->ctx, s->pcgB, b, g.res, g, s->lv[0].labels, s->hostBounds, &s->ioGroups));
    GMG_TRY(pcgDevice(s, s->pcgX, s->pcgB, tol, maxIt, preconditioner, iterations, relResHistory, histCap, histCount));
    const Geom go = ownedGeom(s->lv[0]);
    return downloadValues(s->ctx, x, s->pcgX + int64_t(s->lv[0].ownLo) * g.plane, g.res, go, false, s->hostBounds, &s->ioGroups, s->lv[0].ownLo);
}

extern "C" int gmg_pcg(gmg_solver *s, double *x, const double *b, double tol, int maxIt, int preconditioner, int *iterations, double *relResHistory,
		       int histCap, int *histCount)
{
    return pcgHost(s, x, b, tol, maxIt, preconditioner, iterations, relResHistory, histCap, histCount, false);
}

extern "C" int gmg_pcg_from_zero(gmg_solver *s, double *x, const double *b, double tol, int maxIt, int preconditioner, int *iterations, double *relResHistory,
				 int histCap, int *histCount)
{
    return pcgHost(s, x, b, tol, maxIt, preconditioner, iterations, relResHistory, histCap, histCount, true);
}


// ====================================================================================================
// the steps either side of the solve (gmg_frontend.cuh).  Host arrays in and out, like the builders above.
// ====================================================================================================
namespace
{
struct DevBuf
{
    void *p = nullptr;
    ~DevBuf() { if (p) devFree(p); }
    template <typename T>
    T *as() { return static_cast<T *>(p); }
};
template <typename T>
int devUpload(gmg_ctx *ctx, DevBuf &b, const T *host, int64_t n)
{
    GMG_CUDA(devMalloc(&b.p, sizeof(T) * size_t(std::max<int64_t>(n, 1))));
    if (host) GMG_CUDA(cudaMemcpyAsync(b.p, host, sizeof(T) * size_t(n), cudaMemcpyHostToDevice, ctx->stream));
    return GMG_OK;
}
int64_t cellsOf(const int64_t r[3]) { return r[0] * r[1] * r[2]; }
int64_t facesOf(const int64_t r[3], int axis) { return cellsOf(r) / r[axis] * (r[axis] + 1); }
BaseBox baseBox(const int64_t res[3], const int64_t *expRes, const int64_t *offset)
{
    BaseBox g;
    for (int a = 0; a < 3; ++a) { g.r[a] = res[a]; g.e[a] = expRes ? expRes[a] : res[a]; g.o[a] = offset ? offset[a] : 0; }
    return g;
}
} // namespace

extern "C" int gmg_build_domain_labels(gmg_ctx *ctx, const int32_t *material, const int64_t res[3], int32_t *labels)
{
    if (!ctx || !material || !res || !labels) return invalid("gmg_build_domain_labels: null argument");
    GMG_CUDA(enterCtx(ctx));
    const int64_t n = cellsOf(res);
    DevBuf m, l;
    GMG_TRY(devUpload(ctx, m, material, n));
    GMG_TRY(devUpload<int32_t>(ctx, l, nullptr, n));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_fe_domain_labels<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(l.as<int32_t>(), m.as<int32_t>(), n);
    }
    GMG_CUDA(cudaMemcpyAsync(labels, l.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMG_OK;
}

extern "C" int gmg_build_boundary_weights(gmg_ctx *ctx, const float *cutCell, const float *liquidSurface, const float *validFaces, const int32_t *domainLabels,
					  const int64_t res[3], int axis, double *weights)
{
    if (!ctx || !cutCell || !liquidSurface || !validFaces || !domainLabels || !res || !weights || axis < 0 || axis > 2) return invalid("gmg_build_boundary_weights: bad argument");
    GMG_CUDA(enterCtx(ctx));
    const int64_t n = cellsOf(res), nf = facesOf(res, axis);
    DevBuf cc, ls, vf, dl, w;
    GMG_TRY(devUpload(ctx, cc, cutCell, nf));
    GMG_TRY(devUpload(ctx, ls, liquidSurface, n));
    GMG_TRY(devUpload(ctx, vf, validFaces, nf));
    GMG_TRY(devUpload(ctx, dl, domainLabels, n));
    GMG_TRY(devUpload<double>(ctx, w, nullptr, nf));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_fe_boundary_weights<<<unsigned(divUp(nf, BLOCK)), BLOCK, 0, ctx->stream>>>(w.as<double>(), cc.as<float>(), ls.as<float>(), vf.as<float>(), dl.as<int32_t>(),
										     baseBox(res, nullptr, nullptr), axis);
    }
    GMG_CUDA(cudaMemcpyAsync(weights, w.p, sizeof(double) * nf, cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMG_OK;
}

// rhs / solution: expanded host grids, written on the LIQUID cells only (pass them zero-filled, like the reference's
// rhsGrid.constant(0), GFS.cpp:385): only the base box travels
static int expandedBoxIO(gmg_ctx *ctx, double *host, double *dev, const BaseBox &g, bool toDevice)
{
    cudaMemcpy3DParms p = {};
    const cudaPitchedPtr hp = make_cudaPitchedPtr(host, size_t(g.e[0]) * sizeof(double), size_t(g.e[0]) * sizeof(double), size_t(g.e[1]));
    const cudaPitchedPtr dp = make_cudaPitchedPtr(dev, size_t(g.e[0]) * sizeof(double), size_t(g.e[0]) * sizeof(double), size_t(g.e[1]));
    const cudaPos pos = make_cudaPos(size_t(g.o[0]) * sizeof(double), size_t(g.o[1]), size_t(g.o[2]));
    p.srcPtr = toDevice ? hp : dp;
    p.dstPtr = toDevice ? dp : hp;
    p.srcPos = pos;
    p.dstPos = pos;
    p.extent = make_cudaExtent(size_t(g.r[0]) * sizeof(double), size_t(g.r[1]), size_t(g.r[2]));
    p.kind = toDevice ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    GMG_CUDA(cudaMemcpy3DAsync(&p, ctx->stream));
    return GMG_OK;
}

extern "C" int gmg_build_rhs(gmg_ctx *ctx, const int32_t *material, const float *const velocity[3], const float *const cutCell[3],
			     const float *const solidVelocity[3], const int64_t res[3], const int64_t expRes[3], const int64_t offset[3], double *rhs)
{
    if (!ctx || !material || !velocity || !cutCell || !res || !expRes || !offset || !rhs) return invalid("gmg_build_rhs: null argument");
    GMG_CUDA(enterCtx(ctx));
    const int64_t n = cellsOf(res);
    const BaseBox g = baseBox(res, expRes, offset);
    DevBuf m, r, v[3], c[3], sv[3];
    FeFields fl;
    GMG_TRY(devUpload(ctx, m, material, n));
    for (int a = 0; a < 3; ++a)
    {
	if (!velocity[a] || !cutCell[a]) return invalid("gmg_build_rhs: null field");
	GMG_TRY(devUpload(ctx, v[a], velocity[a], facesOf(res, a)));
	GMG_TRY(devUpload(ctx, c[a], cutCell[a], facesOf(res, a)));
	fl.velocity[a] = v[a].as<float>();
	fl.cutCell[a] = c[a].as<float>();
	fl.solidVelocity[a] = nullptr;
	if (solidVelocity && solidVelocity[a])
	{
	    GMG_TRY(devUpload(ctx, sv[a], solidVelocity[a], facesOf(res, a)));
	    fl.solidVelocity[a] = sv[a].as<float>();
	}
    }
    GMG_TRY(devUpload<double>(ctx, r, nullptr, g.e[0] * g.e[1] * g.e[2]));
    GMG_TRY(expandedBoxIO(ctx, rhs, r.as<double>(), g, true));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_fe_rhs<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(r.as<double>(), m.as<int32_t>(), fl, g);
    }
    GMG_TRY(expandedBoxIO(ctx, rhs, r.as<double>(), g, false));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMG_OK;
}

extern "C" int gmg_apply_old_pressure(gmg_ctx *ctx, const float *pressure, const int32_t *material, const int64_t res[3], const int64_t expRes[3],
				      const int64_t offset[3], double *solution)
{
    if (!ctx || !pressure || !material || !res || !expRes || !offset || !solution) return invalid("gmg_apply_old_pressure: null argument");
    GMG_CUDA(enterCtx(ctx));
    const int64_t n = cellsOf(res);
    const BaseBox g = baseBox(res, expRes, offset);
    DevBuf m, pr, x;
    GMG_TRY(devUpload(ctx, m, material, n));
    GMG_TRY(devUpload(ctx, pr, pressure, n));
    GMG_TRY(devUpload<double>(ctx, x, nullptr, g.e[0] * g.e[1] * g.e[2]));
    GMG_TRY(expandedBoxIO(ctx, solution, x.as<double>(), g, true));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_fe_old_pressure<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(x.as<double>(), pr.as<float>(), m.as<int32_t>(), g);
    }
    GMG_TRY(expandedBoxIO(ctx, solution, x.as<double>(), g, false));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMG_OK;
}

extern "C" int gmg_apply_solution_to_pressure(gmg_ctx *ctx, float *pressure, const int32_t *material, const double *solution, const int64_t res[3],
					      const int64_t expRes[3], const int64_t offset[3])
{
    if (!ctx || !pressure || !material || !solution || !res || !expRes || !offset) return invalid("gmg_apply_solution_to_pressure: null argument");
    GMG_CUDA(enterCtx(ctx));
    const int64_t n = cellsOf(res);
    const BaseBox g = baseBox(res, expRes, offset);
    DevBuf m, pr, x;
    GMG_TRY(devUpload(ctx, m, material, n));
    GMG_TRY(devUpload(ctx, pr, pressure, n));
    GMG_TRY(devUpload<double>(ctx, x, nullptr, g.e[0] * g.e[1] * g.e[2]));
    GMG_TRY(expandedBoxIO(ctx, const_cast<double *>(solution), x.as<double>(), g, true));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_fe_solution_to_pressure<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(pr.as<float>(), x.as<double>(), m.as<int32_t>(), g);
    }
    GMG_CUDA(cudaMemcpyAsync(pressure, pr.p, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMG_OK;
}

extern "C" int gmg_apply_pressure_gradient(gmg_ctx *ctx, float *velocity, const float *liquidSurface, const float *pressure, const float *validFaces,
					   const int32_t *material, const int64_t res[3], int axis)
{
    if (!ctx || !velocity || !liquidSurface || !pressure || !validFaces || !material || !res || axis < 0 || axis > 2) return invalid("gmg_apply_pressure_gradient: bad argument");
    GMG_CUDA(enterCtx(ctx));
    const int64_t n = cellsOf(res), nf = facesOf(res, axis);
    DevBuf v, ls, pr, vf, m;
    GMG_TRY(devUpload(ctx, v, velocity, nf));
    GMG_TRY(devUpload(ctx, ls, liquidSurface, n));
    GMG_TRY(devUpload(ctx, pr, pressure, n));
    GMG_TRY(devUpload(ctx, vf, validFaces, nf));
    GMG_TRY(devUpload(ctx, m, material, n));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_fe_pressure_gradient<<<unsigned(divUp(nf, BLOCK)), BLOCK, 0, ctx->stream>>>(v.as<float>(), ls.as<float>(), pr.as<float>(), vf.as<float>(), m.as<int32_t>(),
										      baseBox(res, nullptr, nullptr), axis);
    }
    GMG_CUDA(cudaMemcpyAsync(velocity, v.p, sizeof(float) * nf, cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMG_OK;
}

extern "C" int gmg_build_material_labels(gmg_ctx *ctx, const float *liquidSurface, const float *solidSurface, const float *const cutCell[3], const int64_t res[3],
					 int32_t *material)
{
    if (!ctx || !liquidSurface || !solidSurface || !cutCell || !res || !material) return invalid("gmg_build_material_labels: null argument");
    GMG_CUDA(enterCtx(ctx));
    const int64_t n = cellsOf(res);
    DevBuf ls, so, c[3], m;
    FeCut fc;
    GMG_TRY(devUpload(ctx, ls, liquidSurface, n));
    GMG_TRY(devUpload(ctx, so, solidSurface, n));
    for (int a = 0; a < 3; ++a)
    {
	if (!cutCell[a]) return invalid("gmg_build_material_labels: null field");
	GMG_TRY(devUpload(ctx, c[a], cutCell[a], facesOf(res, a)));
	fc.cutCell[a] = c[a].as<float>();
    }
    GMG_TRY(devUpload<int32_t>(ctx, m, nullptr, n));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_fe_material_labels<<<unsigned(divUp(n, BLOCK)), BLOCK, 0, ctx->stream>>>(m.as<int32_t>(), ls.as<float>(), so.as<float>(), fc, baseBox(res, nullptr, nullptr));
    }
    GMG_CUDA(cudaMemcpyAsync(material, m.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMG_OK;
}

extern "C" int gmg_build_valid_faces(gmg_ctx *ctx, const int32_t *material, const float *cutCell, const int64_t res[3], int axis, float *validFaces)
{
    if (!ctx || !material || !cutCell || !res || !validFaces || axis < 0 || axis > 2) return invalid("gmg_build_valid_faces: bad argument");
    GMG_CUDA(enterCtx(ctx));
    const int64_t n = cellsOf(res), nf = facesOf(res, axis);
    DevBuf m, c, v;
    GMG_TRY(devUpload(ctx, m, material, n));
    GMG_TRY(devUpload(ctx, c, cutCell, nf));
    GMG_TRY(devUpload<float>(ctx, v, nullptr, nf));
    {
	GMG_LAUNCH(ctx, KC_SETUP, 0);
	k_fe_valid_faces<<<unsigned(divUp(nf, BLOCK)), BLOCK, 0, ctx->stream>>>(v.as<float>(), m.as<int32_t>(), c.as<float>(), baseBox(res, nullptr, nullptr), axis);
    }
    GMG_CUDA(cudaMemcpyAsync(validFaces, v.p, sizeof(float) * nf, cudaMemcpyDeviceToHost, ctx->stream));
    GMG_CUDA(cudaStreamSynchronize(ctx->stream));
    return GMG_OK;
}
