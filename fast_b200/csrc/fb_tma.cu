// fb_tma.cu -- TMA tensor maps of the per-zone SoA arrays (rop levels, xmut) for the k-marching rhs kernel k_rhs3.
//
// An array a(ni,nj,nk,nvar) (i fastest, variable stride = NDIMDX doubles) is described as a rank-4 FP64 tensor; the box
// is (36, rows, 1, box_vars): one k plane of a 31 x RJ tile with its 2-cell halo, all variables at once.  Out-of-range
// box elements are zero-filled by the hardware.  Maps are cached per (pointer, shape): the three time levels rotate
// by pointer swap, so each zone ends up with three rop maps + one xmut map.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <memory>
#include <mutex>

typedef CUresult (*EncodeTiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled_t encode_fn()
{
    static EncodeTiled_t fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiled_t)p;
    }();
    return fn;
}

const CUtensorMap *fb_tma_map(const double *base, int ni, int nj, int nk, long long var_stride, int nvar_total, int box_rows, int box_vars)
{
    typedef std::tuple<int, const double *, int, int, int, long long, int, int, int> Key;
    static std::map<Key, const CUtensorMap *> cache;      // device addresses (never freed: a few 128-byte blocks per zone)
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    int dev = 0; cudaGetDevice(&dev);
    const Key key(dev, base, ni, nj, nk, var_stride, nvar_total, box_rows, box_vars);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::unique_ptr<CUtensorMap> m;
    EncodeTiled_t enc = encode_fn();
    // global strides must be multiples of 16 bytes, the base 16-byte aligned
    if (enc && base && (ni % 2) == 0 && (var_stride % 2) == 0 && ((uintptr_t)base % 16) == 0) {
        m.reset(new CUtensorMap);
        const cuuint64_t dims[4] = {(cuuint64_t)ni, (cuuint64_t)nj, (cuuint64_t)nk, (cuuint64_t)nvar_total};
        const cuuint64_t strides[3] = {(cuuint64_t)ni * 8, (cuuint64_t)ni * nj * 8, (cuuint64_t)var_stride * 8};
        const cuuint32_t box[4] = {36, (cuuint32_t)box_rows, 1, (cuuint32_t)box_vars};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        // a 64-bit element type of the right size; the data are moved, never interpreted
        const CUresult r = enc(m.get(), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)base, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            static bool warned = false;
            if (!warned) { fprintf(stderr, "fast_b200: cuTensorMapEncodeTiled failed (%d): rhs falls back to the L1 path\n", (int)r); warned = true; }
            m.reset();
        }
    }
    // The kernels read the descriptor from DEVICE memory, one address per (array, shape): a descriptor handed over as a
    // __grid_constant__ kernel parameter sits at the same parameter-bank address in every launch, and with the per-zone rhs
    // kernels of a phase running concurrently from several streams tiles were occasionally staged through another launch's
    // descriptor (profiles/r02_race.md: nondeterministic 1e-5 perturbations, gone with unique descriptor addresses).
    const CUtensorMap *out = nullptr;
    if (m) {
        CUtensorMap *d = nullptr;
        if (cudaMalloc(&d, sizeof(CUtensorMap)) == cudaSuccess && cudaMemcpy(d, m.get(), sizeof(CUtensorMap), cudaMemcpyHostToDevice) == cudaSuccess) out = d;
        else cudaGetLastError();
    }
    cache[key] = out;
    return out;
}

// Skewed view of a zone array a(ni,nj,nk,nvar) for the LU sweep kernel k_lu_cell (fb_lu.cu): element (c0, c1, c2, e) ->
// a[e][c0 + c1 (ni - 2) + c2 (ni nj - 2)], i.e. c0 = i + 2 j + 2 k -- the granules (2 cells along i) that a tile of lines
// needs at one super-step of the i+j+k wavefront form a dense (2, 8, box_k, box_vars) box.  rank3: k-invariant plane arrays
// (3dhomo normals), c0 = i + 2 j, box (2, 8, box_vars).  The view overlaps itself (stride of c1 < extent of c0); the largest
// address it can produce is the last cell of the array, out-of-range coordinates are zero-filled (tools/probe/tma_skew_probe.cu).
bool fb_tma_encode_skew(CUtensorMap *out, const double *base, int ni, int nj, int nk, long long estride, int nvar, int box_k, int box_vars, int rank3)
{
    EncodeTiled_t enc = encode_fn();
    if (!enc || !base || (ni % 2) || (estride % 2) || ((uintptr_t)base % 16)) return false;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r;
    if (!rank3) {
        const cuuint64_t dims[4] = {(cuuint64_t)(ni + 2 * nj + 2 * nk), (cuuint64_t)nj, (cuuint64_t)nk, (cuuint64_t)nvar};
        const cuuint64_t strides[3] = {(cuuint64_t)(ni - 2) * 8, ((cuuint64_t)ni * nj - 2) * 8, (cuuint64_t)estride * 8};
        const cuuint32_t box[4] = {2, 8, (cuuint32_t)box_k, (cuuint32_t)box_vars};
        r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[3] = {(cuuint64_t)(ni + 2 * nj), (cuuint64_t)nj, (cuuint64_t)nvar};
        const cuuint64_t strides[2] = {(cuuint64_t)(ni - 2) * 8, (cuuint64_t)estride * 8};
        const cuuint32_t box[3] = {2, 8, (cuuint32_t)box_vars};
        r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS;
}

// launch counters of the two rhs paths (fb_counter in the C ABI; bench.py names the kernel it times from them)
long long g_fb_rhs_launches[2] = {0, 0};   // [0] k_rhs3 (TMA ring), [1] k_rhs2 (L1 path)
