// Probe of TMA facts this library relies on (run on a B200; each case in its own process because a
// faulting case kills the context).  usage: tma_probe <case>
//   0: map in kernel params at a small offset      1: map at a 9 KB offset inside a 10 KB param struct
//   2: map in global memory                         3: permuted dims (x, channel, y) at small offset
// Also prints the host cost of cuTensorMapEncodeTiled.
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct alignas(64) Small { CUtensorMap m; };
struct alignas(64) Big { CUtensorMap m[80]; };

__device__ int g_mode;   // bit 0: no TMA (plain arrive), bit 1: omit the .tile qualifier
__device__ void run(const void* map, float* out, int x, int c1, int c2, int n) {
    __shared__ __align__(128) float buf[4 * 8 * 32];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_sa = (unsigned)__cvta_generic_to_shared(&bar);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_sa) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (g_mode & 1) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_sa), "r"(0) : "memory");
        } else {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_sa), "r"(n * 4) : "memory");
            if (g_mode & 2)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(dst), "l"(map), "r"(x), "r"(c1), "r"(c2), "r"(bar_sa) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(dst), "l"(map), "r"(x), "r"(c1), "r"(c2), "r"(bar_sa) : "memory");
        }
    }
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(bar_sa), "r"(0) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
__global__ void k_small(const __grid_constant__ Small s, float* out, int x, int c1, int c2, int n) { run(&s.m, out, x, c1, c2, n); }
__global__ void k_big(const __grid_constant__ Big b, int idx, float* out, int x, int c1, int c2, int n) { run(&b.m[idx], out, x, c1, c2, n); }
__global__ void k_glob(const CUtensorMap* g, float* out, int x, int c1, int c2, int n) { run(g, out, x, c1, c2, n); }

int main(int argc, char** argv) {
    const int cs = argc > 1 ? atoi(argv[1]) : 0;
    const int mode = argc > 2 ? atoi(argv[2]) : 0;
    const int xarg = argc > 3 ? atoi(argv[3]) : 61;
    cudaMemcpyToSymbol(g_mode, &mode, sizeof(int));
    const int N = 2, C = 8, H = 40, W = 84, BW = 24, BH = 8, BC = 4;
    std::vector<float> h((size_t)N * C * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, 4096 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)f;
    if (!enc) { printf("no encode fn\n"); return 2; }
    const bool perm = cs == 3 || cs == 4;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], ones[3] = {1, 1, 1};
    dims[0] = W; box[0] = BW;
    if (perm) { dims[1] = (cuuint64_t)N * C; dims[2] = H; strides[0] = (cuuint64_t)H * W * 4; strides[1] = W * 4; box[1] = BC; box[2] = BH; }
    else { dims[1] = H; dims[2] = (cuuint64_t)N * C; strides[0] = W * 4; strides[1] = (cuuint64_t)H * W * 4; box[1] = BH; box[2] = BC; }
    CUtensorMap m;
    auto t0 = std::chrono::steady_clock::now();
    CUresult r = CUDA_SUCCESS;
    for (int i = 0; i < 1000; ++i)
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    auto t1 = std::chrono::steady_clock::now();
    printf("case %d encode rc=%d  %.3f us per encode\n", cs, (int)r, std::chrono::duration<double, std::micro>(t1 - t0).count() / 1000);
    if (r != CUDA_SUCCESS) return 3;
    const int n = BW * BH * BC;
    // box at x = 61 (unaligned, runs 1 column past W = 84 -> zero fill), channel 5 (image 0), y = 35 (3 rows past H)
    const int x = xarg, ch = 5, y = 35;
    if (cs == 0 || cs == 3) { Small s; s.m = m; k_small<<<1, 128>>>(s, out, x, perm ? ch : y, perm ? y : ch, n); }
    else if (cs == 1 || cs == 4) { static Big b; b.m[72] = m; k_big<<<1, 128>>>(b, 72, out, x, perm ? ch : y, perm ? y : ch, n); }
    else { CUtensorMap* g; cudaMalloc(&g, sizeof(m)); cudaMemcpy(g, &m, sizeof(m), cudaMemcpyHostToDevice); k_glob<<<1, 128>>>(g, out, x, y, ch, n); }
    cudaError_t e = cudaDeviceSynchronize();
    printf("case %d mode %d x %d kernel: %s\n", cs, mode, x, cudaGetErrorString(e));
    if (e != cudaSuccess) return 4;
    std::vector<float> o(n);
    cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < BC; ++c)
        for (int rr = 0; rr < BH; ++rr)
            for (int xx = 0; xx < BW; ++xx) {
                const int gy = y + rr, gx = x + xx, gc = ch + c;
                const float want = (gy < H && gx < W && gc < N * C) ? h[((size_t)gc * H + gy) * W + gx] : 0.0f;
                const float got = perm ? o[(rr * BC + c) * BW + xx] : o[(c * BH + rr) * BW + xx];
                if (want != got) ++bad;
            }
    printf("case %d mismatches %d of %d (layout %s)\n", cs, bad, n, perm ? "[row][ch][BW]" : "[ch][row][BW]");
    return bad ? 5 : 0;
}
