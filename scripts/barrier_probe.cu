// Micro-benchmark behind the persistent-cycle design: cost of one dependent step as (a) a kernel boundary inside a CUDA graph
// with programmatic dependent launch, (b) the same without PDL, (c) a grid-wide barrier inside one cooperative kernel.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/barrier_probe scripts/barrier_probe.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_step(double *p, int n)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] * 1.0000001 + 1.0;
}

__device__ __forceinline__ void gridBarrier(unsigned *counter, unsigned &generation)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
	++generation;
	const unsigned target = generation * gridDim.x;
	__threadfence();
	atomicAdd(counter, 1u);
	unsigned v;
	do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
    }
    __syncthreads();
}

__global__ void k_persistent(double *p, int n, int steps, unsigned *counter, int useCg)
{
    unsigned gen = 0;
    cg::grid_group g = cg::this_grid();
    for (int s = 0; s < steps; ++s)
    {
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = p[i] * 1.0000001 + 1.0;
	if (useCg) g.sync();
	else gridBarrier(counter, gen);
    }
}

int main()
{
    const int steps = 200;
    double *p;
    unsigned *counter;
    CK(cudaMalloc(&p, sizeof(double) << 22));
    CK(cudaMemset(p, 0, sizeof(double) << 22));
    CK(cudaMalloc(&counter, 4));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int n : {8192, 65536, 524288})
    {
	for (int pdl = 0; pdl < 2; ++pdl)
	{
	    cudaGraph_t graph;
	    cudaGraphExec_t exec;
	    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
	    for (int s = 0; s < steps; ++s)
	    {
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3((n + 255) / 256);
		cfg.blockDim = dim3(256);
		cfg.stream = st;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr;
		cfg.numAttrs = pdl;
		CK(cudaLaunchKernelEx(&cfg, k_step, p, n));
	    }
	    CK(cudaStreamEndCapture(st, &graph));
	    CK(cudaGraphInstantiate(&exec, graph, 0));
	    for (int r = 0; r < 3; ++r) CK(cudaGraphLaunch(exec, st));
	    CK(cudaEventRecord(e0, st));
	    for (int r = 0; r < 10; ++r) CK(cudaGraphLaunch(exec, st));
	    CK(cudaEventRecord(e1, st));
	    CK(cudaStreamSynchronize(st));
	    float ms;
	    cudaEventElapsedTime(&ms, e0, e1);
	    printf("n=%7d graph chain pdl=%d: %.2f us per step\n", n, pdl, ms * 1e3 / (10 * steps));
	    cudaGraphExecDestroy(exec);
	    cudaGraphDestroy(graph);
	}
	for (int cfgI = 0; cfgI < 4; ++cfgI)
	{
	    const int threads = (cfgI & 1) ? 1024 : 512;
	    const int perSm = (cfgI & 1) ? 1 : 2;
	    const int useCg = cfgI >> 1;
	    int grid = 148 * perSm;
	    CK(cudaMemsetAsync(counter, 0, 4, st));
	    int nn = n, stp = steps;
	    void *args[] = {&p, &nn, &stp, &counter, (void *)&useCg};
	    CK(cudaLaunchCooperativeKernel((void *)k_persistent, dim3(grid), dim3(threads), args, 0, st));
	    CK(cudaMemsetAsync(counter, 0, 4, st));
	    CK(cudaEventRecord(e0, st));
	    CK(cudaLaunchCooperativeKernel((void *)k_persistent, dim3(grid), dim3(threads), args, 0, st));
	    CK(cudaEventRecord(e1, st));
	    CK(cudaStreamSynchronize(st));
	    float ms;
	    cudaEventElapsedTime(&ms, e0, e1);
	    printf("n=%7d persistent %d x %d %s: %.2f us per step\n", n, grid, threads, useCg ? "cg::grid.sync" : "own barrier", ms * 1e3 / steps);
	}
    }
    return 0;
}
