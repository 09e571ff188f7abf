// probe: does cuTensorMapEncodeTiled accept a "skewed" overlapping view (stride of dim 1 < extent of dim 0), do negative / out-of-range
// start coordinates zero-fill, and does an odd (8-byte aligned, not 16-byte aligned) box start work?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
typedef CUresult (*EncodeTiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, double *out)
{
    __shared__ __align__(128) double buf[2 * 8 * 4 * 3];
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"((unsigned)sizeof(buf)));
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"((unsigned)__cvta_generic_to_shared(buf)), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(0), "r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
    }
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < 2 * 8 * 4 * 3; i += blockDim.x) out[i] = buf[i];
}
int main()
{
    const int ni = 20, nj = 12, nk = 6, nvar = 3, nd = ni * nj * nk;
    std::vector<double> h((size_t)nd * nvar);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
    double *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, 192 * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiled_t enc = (EncodeTiled_t)p;
    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)(ni + 2 * nj + 2 * nk), (cuuint64_t)nj, (cuuint64_t)nk, (cuuint64_t)nvar};
    const cuuint64_t strides[3] = {(cuuint64_t)(ni - 2) * 8, (cuuint64_t)(ni * nj - 2) * 8, (cuuint64_t)nd * 8};
    const cuuint32_t box[4] = {2, 8, 4, 3}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode skewed view: %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    std::vector<double> g(192);
    const int tests[4][3] = {{10, 1, 1}, {-2, 0, 0}, {6, -3, 1}, {40, 8, 4}};
    for (int t = 0; t < 4; t++) {
        const int c0 = tests[t][0], c1 = tests[t][1], c2 = tests[t][2];
        k<<<1, 64>>>(tm, c0, c1, c2, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test %d (%d,%d,%d): launch error %s\n", t, c0, c1, c2, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(g.data(), o, 192 * 8, cudaMemcpyDeviceToHost);
        int bad = 0, zero = 0;
        for (int e2 = 0; e2 < 3; e2++) for (int kk = 0; kk < 4; kk++) for (int jj = 0; jj < 8; jj++) for (int c = 0; c < 2; c++) {
            const int a0 = c0 + c, a1 = c1 + jj, a2 = c2 + kk;
            const bool in = a0 >= 0 && a0 < (int)dims[0] && a1 >= 0 && a1 < nj && a2 >= 0 && a2 < nk;
            const double want = in ? (double)((long)e2 * nd + a0 + (long)a1 * (ni - 2) + (long)a2 * (ni * nj - 2)) : 0.;
            const double got = g[((e2 * 4 + kk) * 8 + jj) * 2 + c];
            if (got != want) bad++;
            if (!in) zero++;
        }
        printf("test %d start (%d,%d,%d): mismatches %d (out-of-range elements %d)\n", t, c0, c1, c2, bad, zero);
    }
    return 0;
}
