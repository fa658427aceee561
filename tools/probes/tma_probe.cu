// Development probe: one cp.async.bulk.tensor.3d box load with the given box / coordinates; prints the checksum or the CUDA error.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; ./tma_probe bw bh x y plane [align]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int bw, int bh, int x, int y, int pl, int off, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* region = reinterpret_cast<float*>(sm + off);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + off + ((size_t)bw * bh * 4 + 127) / 128 * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bw * bh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(region)),
                     "l"(&tm), "r"(s32(bar)), "r"(x), "r"(y), "r"(pl)
                     : "memory");
    }
    __syncthreads();
    uint32_t ok = 0, spins = 0;
    while (!ok && spins < (1u << 22)) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(bar)) : "memory");
        ++spins;
    }
    float s = 0.f;
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) s += region[i];
    atomicAdd(out, ok ? s : 1e30f);
}
int main(int argc, char** argv) {
    const int bw = atoi(argv[1]), bh = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]), pl = atoi(argv[5]), off = argc > 6 ? atoi(argv[6]) : 0;
    const int W = 512, H = 512, P = 6;
    float* d; cudaMalloc(&d, (size_t)W * H * P * 4);
    float* h = (float*)malloc((size_t)W * H * P * 4);
    for (size_t i = 0; i < (size_t)W * H * P; ++i) h[i] = 1.0f;
    cudaMemcpy(d, h, (size_t)W * H * P * 4, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm;
    const cuuint64_t dims[3] = {W, H, P}; const cuuint64_t strides[2] = {W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const size_t smem = off + ((size_t)bw * bh * 4 + 127) / 128 * 128 + 16;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(tm, bw, bh, x, y, pl, off, out);
    cudaError_t e = cudaDeviceSynchronize();
    float res = -1; cudaMemcpy(&res, out, 4, cudaMemcpyDeviceToHost);
    printf("box %dx%d at (%d,%d,%d) off %d: encode %d, %s, in-bounds elements summed = %.0f\n", bw, bh, x, y, pl, off, (int)r, cudaGetErrorString(e), res);
    return 0;
}
