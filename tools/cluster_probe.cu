// How many CTAs of the search kernel's shape can be co-resident for cluster sizes 1, 2, 4, 8 on this GPU?
// (cudaOccupancyMaxActiveClusters with the kernel's 232448 B of dynamic shared memory, 320 threads per CTA.)
// Evidence for DESIGN.md section 4.1: 148 SMs pack perfectly into clusters of 2 (one TPC each); larger clusters must
// fit inside a GPC and strand SMs, so sharing codebook stages across 4 or 8 CTAs by TMA multicast costs more tensor
// throughput than the saved L2->SM traffic could return.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/cluster_probe tools/cluster_probe.cu && tools/cluster_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(320, 1) shape_kernel(float* p) {
  extern __shared__ float smem[];
  if (p) p[threadIdx.x] = smem[threadIdx.x];
}

int main() {
  const int smem = 232448;
  cudaFuncSetAttribute(shape_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(shape_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("device: %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 1024);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, shape_kernel, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d co-resident CTAs (%s)\n", cs, n, n * cs,
           e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  }
  return 0;
}
