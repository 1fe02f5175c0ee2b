// Stand-alone HBM microbenchmark used to calibrate the roofline denominators of the write-dominated
// kernels (RoIAlign forward, paste): how fast can this part *write*, by which store flavour?
// Not part of the product; built by tools/Makefile, run under gpurun, results in profiles/.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int MODE>  // 0 default, 1 .cs, 2 .wt, 3 v8 (256-bit)
__global__ void fill_kernel(uint4* __restrict__ p, size_t n16) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (MODE == 3) {
        for (size_t j = i; j * 2 + 1 < n16; j += stride) {
            asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + 2 * j), "r"(0) : "memory");
        }
        return;
    }
    for (; i < n16; i += stride) {
        if (MODE == 0) p[i] = z;
        if (MODE == 1) __stcs(p + i, z);
        if (MODE == 2) __stwt(p + i, z);
    }
}

// each CTA owns contiguous 16 KB blocks (like a tile writer), 4 x 16 B per thread
template <int MODE>
__global__ void fill_tiles(uint4* __restrict__ p, size_t n16) {
    const size_t tiles = n16 / 1024;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        uint4* q = p + t * 1024;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (MODE == 0) q[threadIdx.x + 256 * k] = z;
            else __stcs(q + threadIdx.x + 256 * k, z);
        }
    }
}

// TMA bulk store: one elected thread streams a zeroed shared buffer to global memory
template <int BYTES>
__global__ void fill_bulk(unsigned char* __restrict__ p, size_t nbytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t tiles = nbytes / BYTES;
        const unsigned s = (unsigned)__cvta_generic_to_shared(sm);
        int inflight = 0;
        for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + t * BYTES), "r"(s), "r"(BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void read_kernel(const uint4* __restrict__ p, size_t n16, unsigned* sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
        acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void copy_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n16) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 x0 = __ldcs(a + i), x1 = __ldcs(a + i + stride), x2 = __ldcs(a + i + 2 * stride), x3 = __ldcs(a + i + 3 * stride);
        __stcs(b + i, x0); __stcs(b + i + stride, x1); __stcs(b + i + 2 * stride, x2); __stcs(b + i + 3 * stride, x3);
    }
}

// one RED.ADD.F32 per element / one vector red per 4 elements, every address touched once
template <int VEC>
__global__ void red_kernel(float* __restrict__ p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (VEC == 1) {
        for (; i < n; i += stride) atomicAdd(p + i, 1.0f);
    } else {
        for (; i * 4 + 3 < n; i += stride)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p + 4 * i), "f"(1.0f) : "memory");
    }
}

// RoIAlign-forward-like write pattern: each warp (MODE 0) or each CTA (MODE 1: warp w takes rows
// w, w+8, ...) streams whole [112][112] fp32 planes, 28 lanes x 16 B per row, persistent CTAs.
template <int MODE>
__global__ void plane_writer(float4* __restrict__ p, size_t nplanes) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float4 z = make_float4(1.f, 2.f, 3.f, 4.f);
    if (MODE == 0) {
        for (size_t pl = (size_t)blockIdx.x * nw + warp; pl < nplanes; pl += (size_t)gridDim.x * nw) {
            float4* q = p + pl * (112 * 28) + lane;
            if (lane < 28)
                for (int r = 0; r < 112; ++r) __stcs(q + r * 28, z);
        }
    } else {
        for (size_t pl = blockIdx.x; pl < nplanes; pl += gridDim.x) {
            float4* q = p + pl * (112 * 28) + lane;
            if (lane < 28)
                for (int r = warp; r < 112; r += nw) __stcs(q + r * 28, z);
        }
    }
}

template <class F>
static double time_ms(F f, int reps = 10) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, tot = 0;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best; tot += ms;
    }
    CK(cudaGetLastError());
    printf("    best %.4f ms  mean %.4f ms\n", best, tot / reps);
    return best;
}

int main(int argc, char** argv) {
    size_t bytes = (argc > 1 ? (size_t)atoll(argv[1]) : 2048) << 20;  // MiB
    unsigned char *a, *b; unsigned* sink;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    const size_t n16 = bytes / 16;
    const double gb = bytes / 1e9;
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("buffer %.2f GB, %d SMs\n", gb, sms);
#define REPORT(name, expr, traffic_gb) { printf("%s\n", name); double ms = time_ms([&] { expr; }); printf("    -> %.1f GB/s\n", (traffic_gb) / ms * 1e3); }
    REPORT("cudaMemsetAsync", CK(cudaMemsetAsync(a, 0, bytes)), gb);
    for (int g : {sms * 2, sms * 4, sms * 8, sms * 16, sms * 64}) {
        char nm[96];
        snprintf(nm, 96, "fill st.v4 default, grid %d x 256", g); REPORT(nm, (fill_kernel<0><<<g, 256>>>((uint4*)a, n16)), gb);
        snprintf(nm, 96, "fill st.v4 .cs, grid %d x 256", g);     REPORT(nm, (fill_kernel<1><<<g, 256>>>((uint4*)a, n16)), gb);
    }
    REPORT("fill st.v4 .wt, grid 8/SM", (fill_kernel<2><<<sms * 8, 256>>>((uint4*)a, n16)), gb);
    REPORT("fill st.v8 (256-bit), grid 8/SM", (fill_kernel<3><<<sms * 8, 256>>>((uint4*)a, n16)), gb);
    REPORT("fill st.v4 one-shot grid (1 x 16B / thread)", (fill_kernel<0><<<(unsigned)(n16 / 256), 256>>>((uint4*)a, n16)), gb);
    REPORT("fill tiles 16KB/CTA-iter default, grid 8/SM", (fill_tiles<0><<<sms * 8, 256>>>((uint4*)a, n16)), gb);
    REPORT("fill tiles 16KB/CTA-iter .cs, grid 8/SM", (fill_tiles<1><<<sms * 8, 256>>>((uint4*)a, n16)), gb);
    REPORT("fill tiles 16KB/CTA .cs, one tile per CTA", (fill_tiles<1><<<(unsigned)(n16 / 1024), 256>>>((uint4*)a, n16)), gb);
    CK(cudaFuncSetAttribute(fill_bulk<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    CK(cudaFuncSetAttribute(fill_bulk<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    REPORT("fill TMA bulk 16KB, 1 CTA/SM", (fill_bulk<16384><<<sms, 128, 16384>>>(a, bytes)), gb);
    REPORT("fill TMA bulk 16KB, 4 CTA/SM", (fill_bulk<16384><<<sms * 4, 128, 16384>>>(a, bytes)), gb);
    REPORT("fill TMA bulk 64KB, 2 CTA/SM", (fill_bulk<65536><<<sms * 2, 128, 65536>>>(a, bytes)), gb);
    for (int g : {sms * 4, sms * 8, sms * 32}) {
        char nm[96];
        snprintf(nm, 96, "read ld.v4 .cs x4, grid %d x 256", g); REPORT(nm, (read_kernel<<<g, 256>>>((const uint4*)a, n16, sink)), gb);
    }
    for (int g : {sms * 4, sms * 8, sms * 32}) {
        char nm[96];
        snprintf(nm, 96, "copy ld/st .cs x4, grid %d x 256 (read+write bytes)", g); REPORT(nm, (copy_kernel<<<g, 256>>>((const uint4*)a, (uint4*)b, n16)), 2 * gb);
    }
    REPORT("cudaMemcpyAsync D2D (read+write bytes)", CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)), 2 * gb);
    REPORT("red.add.f32 scalar, every address once (bytes = 4/elem)", (red_kernel<1><<<sms * 16, 256>>>((float*)a, bytes / 4)), gb);
    REPORT("red.add.v4.f32, every address once", (red_kernel<4><<<sms * 16, 256>>>((float*)a, bytes / 4)), gb);
    {
        const size_t nplanes = bytes / (112 * 112 * 4);
        const double pgb = nplanes * 112.0 * 112 * 4 / 1e9;
        for (int cps : {2, 3, 4, 8}) {
            char nm[96];
            snprintf(nm, 96, "plane writer, warp per plane, %d CTAs/SM x 256", cps);
            REPORT(nm, (plane_writer<0><<<sms * cps, 256>>>((float4*)a, nplanes)), pgb);
            snprintf(nm, 96, "plane writer, CTA per plane, %d CTAs/SM x 256", cps);
            REPORT(nm, (plane_writer<1><<<sms * cps, 256>>>((float4*)a, nplanes)), pgb);
        }
    }
    // smaller-than-L2 fill for reference
    REPORT("fill st.v4 .cs 64 MiB (L2 resident)", (fill_kernel<1><<<sms * 8, 256>>>((uint4*)a, (64u << 20) / 16)), (64u << 20) / 1e9);
    return 0;
}
