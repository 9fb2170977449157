// cluster_probe.cu -- what does one step of a thread-block-cluster kernel cost on B200?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cluster_probe cluster_probe.cu && ./cluster_probe
// Variants of the barrier between steps (16 or 8 CTAs x 1024 or 256 threads, 2000 steps, one DSMEM exchange per step with a
// correctness check of the value read from the next CTA):
//   0 cg::this_cluster().sync()                      (arrive.release + wait.acquire)
//   1 barrier.cluster.arrive.relaxed + wait          (no memory ordering by the barrier itself)
//   2 fence.acq_rel.cluster + relaxed barrier
//   3 __threadfence_block + relaxed barrier
//   4 __syncthreads only                             (baseline, no cross-CTA exchange)
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int MODE>
__device__ __forceinline__ void stepBarrier(cg::cluster_group &cl)
{
    if (MODE == 0) cl.sync();
    else if (MODE == 1) { asm volatile("barrier.cluster.arrive.relaxed.aligned;\n barrier.cluster.wait.aligned;" ::: "memory"); }
    else if (MODE == 2) { asm volatile("fence.acq_rel.cluster;\n barrier.cluster.arrive.relaxed.aligned;\n barrier.cluster.wait.aligned;" ::: "memory"); }
    else if (MODE == 3) { __threadfence_block(); asm volatile("barrier.cluster.arrive.relaxed.aligned;\n barrier.cluster.wait.aligned;" ::: "memory"); }
    else __syncthreads();
}

template <int MODE>
__global__ void k_probe(int steps, int *mismatch, double *sink)
{
    extern __shared__ double sm[];
    cg::cluster_group cl = cg::this_cluster();
    const int rank = cl.block_rank(), n = cl.num_blocks();
    double *a = sm, *b = sm + blockDim.x;
    a[threadIdx.x] = 0.0;
    b[threadIdx.x] = 0.0;
    cl.sync();
    int bad = 0;
    double acc = 0.0;
    for (int s = 1; s <= steps; ++s)
    {
	double *w = (s & 1) ? a : b;
	w[threadIdx.x] = double(s) + rank * 1e-3;
	stepBarrier<MODE>(cl);
	if (MODE != 4)
	{
	    const int peer = (rank + 1) % n;
	    const double v = *cl.map_shared_rank(w + threadIdx.x, peer);
	    if (v != double(s) + peer * 1e-3) ++bad;
	    acc += v;
	}
	else acc += w[(threadIdx.x + 1) % blockDim.x];
    }
    cl.sync();
    if (bad) atomicAdd(mismatch, bad);
    if (acc == 12345.678) *sink = acc;
}

template <int MODE>
void run(int clusterSize, int threads, int steps)
{
    int *mismatch;
    double *sink;
    cudaMalloc(&mismatch, sizeof(int));
    cudaMalloc(&sink, sizeof(double));
    cudaMemset(mismatch, 0, sizeof(int));
    cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusterSize);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = 2 * threads * sizeof(double);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = clusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t err = cudaLaunchKernelEx(&cfg, k_probe<MODE>, 10, mismatch, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    err = cudaLaunchKernelEx(&cfg, k_probe<MODE>, steps, mismatch, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int h = 0;
    cudaMemcpy(&h, mismatch, sizeof(int), cudaMemcpyDeviceToHost);
    printf("mode %d cluster %2d x %4d threads: %.3f us per step, mismatches %d%s\n", MODE, clusterSize, threads, ms * 1e3 / steps, h,
	   err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(mismatch);
    cudaFree(sink);
}

int main()
{
    const int steps = 2000;
    for (int cs : {16, 8, 2})
	for (int th : {1024, 256})
	{
	    run<0>(cs, th, steps);
	    run<1>(cs, th, steps);
	    run<2>(cs, th, steps);
	    run<3>(cs, th, steps);
	    run<4>(cs, th, steps);
	}
    return 0;
}
