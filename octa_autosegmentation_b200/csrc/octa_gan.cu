// f-3: GAN contrast adaptation (resnetGenerator9 inference) for sm_100a.
//
// Replaces, for config #5 (graphs -> 304^2 raster -> GAN -> G_<name>.png), the generator forward of the reference:
//   test.py:58-90 -> GanSegModel.inference (models/gan_seg_model.py:65-79) -> ResnetGenerator.forward
//   (models/networks.py:350-443; resnetGenerator9 = ngf 64, InstanceNorm2d(affine=False), 9 blocks, :502-503)
// and the input transforms of docker/trained_models/GAN/config.yml:49-92 that touch pixels:
//   ScaleIntensityd(0,1) on raster and background, Rotate90d(k=1)+Flipd(0) on the background (= a transpose),
//   AddRandomBackgroundNoised (data/data_transforms.py:498-516): img = max(img, background * U(0,1)) with the speckle
//   drawn from the legacy numpy MT19937 stream, CastToTyped(float32); output writer utils/visualizer.py:338.
//
// Layout.  Activations are bf16, channels-last, stored WITH their spatial padding: [N][H+2P][W+2P][C].  With the padding
// materialised a 3x3 convolution is nine SHIFTED GEMMs over the flat pixel index q = (n*(H+2) + h)*(W+2) + w:
//     raw[q][co] = sum_{kh,kw,ci} act[q + kh*(W+2) + kw][ci] * Wt[co][kh][kw][ci]
// (rows with h >= H or w >= W are computed too -- 1.3..5 % extra -- and never read).  So every A tile of the implicit GEMM
// is a plain 2-D TMA box [128 pixels x 64 channels] at a shifted row coordinate; no im2col buffer exists.
//
// k_gan_conv3 is the tensor-core kernel: TMA (128B swizzle) -> shared-memory ring -> tcgen05.mma (cta_group::1,
// kind::f16, bf16 x bf16 -> fp32, M=128, N=Cout=64/128/256, K=16) with the accumulator in TMEM -> tcgen05.ld epilogue -> bf16 rows.
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane), warps 2..5 = epilogue
// (one TMEM lane quarter each).  The kernel is persistent (one CTA per SM walks the M tiles) with TWO accumulators in TMEM, so
// the epilogue of tile i runs under the main loop of tile i+1, and the TMA ring (192 KB) runs ahead across tiles.  Conv biases that are followed by InstanceNorm(affine=False) cancel exactly in the
// mean subtraction and are not applied.
// Instance-norm statistics are taken in the conv epilogue (per-tile partial sums from the fp32 accumulators, finalised by a tiny
// kernel, deterministic).  The remaining layers are bandwidth-bound element kernels: norm+ReLU(+skip) into the next
// padded buffer (zero or reflect border), blur-pool down / bilinear up (with the norm applied per tap).  The 7x7 stem (1 -> 64)
// and head (64 -> 1) are GEMMs on the same kernel: im2col rows of 49 taps, resp. per-pixel tap responses summed afterwards.
#include "octa_common.h"
#include "octa_rng.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

namespace {

using bf16 = __nv_bfloat16;

// ------------------------------------------------------------------------------------------
// small PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a lost arrival traps (the launch fails with an error) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of this warp's TMEM lane quarter
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor: K-major tile of [rows][64 bf16] (128-byte rows), 128B swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------
// 3x3 convolution as nine shifted GEMMs on tcgen05
// ------------------------------------------------------------------------------------------
constexpr int CONV_THREADS = 192;
template <int BLOCK_N, int CTAS> struct ConvCfg {                          // CTAS = resident CTAs per SM (1 or 2)
    static constexpr uint32_t A_BYTES = 128 * 128, B_BYTES = BLOCK_N * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (192 * 1024 / CTAS) / (int)STAGE_BYTES;  // 1 CTA/SM: N = 64: 8 x 24 KB, 128: 6 x 32 KB, 256: 4 x 48 KB
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;
    static constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;                     // two accumulators: 128 / 256 / 512 columns
    static_assert(TMEM_COLS * CTAS <= 512 && STAGES >= 2, "TMEM / shared memory budget");
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// a[j] of lane l = value of row l, column j  ->  returns the sum over the warp's 32 rows of column `lane`.  Five butterfly
// steps; each halves the columns a lane still carries (31 shuffles, fixed summation order: deterministic).
__device__ __forceinline__ float warp_colsum(float (&a)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? a[i] : a[i + off];
            const float keep = up ? a[i + off] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return a[0];
}

// Instance-norm statistics fused into the conv epilogue (EPI 0, st.part != nullptr): every M tile leaves the per-channel sum and
// sum of squares of its valid rows (h < H, w < W of the padded grid) in part[tile][channel]; rows of a tile that already belong
// to the next image go to carry[image][quarter][channel].  k_gan_stats_tiles adds a fixed sequence of partials per image.
struct ConvStats {
    float2* part;       // [num_tiles][cout]
    float2* carry;      // [n_images][4][cout]
    int hpwp, H, W;     // rows per image of the padded grid, valid extent
};

// Persistent, warp-specialised: one CTA per SM walks the M tiles (128 flat pixels x all BLOCK_N = Cout channels).
//   warp 0 (one lane)  TMA producer: ring of STAGES x (A 128x64 + B BLOCK_Nx64) bf16 tiles, 128B swizzle, runs ahead across tiles
//   warp 1 (one lane)  MMA issuer: 4 x tcgen05.mma (K = 16) per stage into accumulator (tile & 1) of TMEM; tcgen05.commit frees
//                      the stage, and after the last k-block hands the accumulator to the epilogue
//   warps 2..5         epilogue of the PREVIOUS tile while the next one accumulates: tcgen05.ld (one TMEM lane quarter each)
// EPI 0: bf16 rows out[q][cout] (the conv layers).  EPI 1: fp32 planes out[c][q], c < cout (the 7x7 head: one GEMM gives the
// response of every input pixel to every tap, k_gan_head_sum then adds 49 shifted planes -- all loads coalesced).
template <int BLOCK_N, int EPI, int CTAS>
__global__ void __launch_bounds__(CONV_THREADS, CTAS)
k_gan_conv3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, void* __restrict__ out,
            int m_total, int wp, int cin_blocks, int cout, int taps, int kwn, int num_tiles, ConvStats st) {
    using Cfg = ConvCfg<BLOCK_N, CTAS>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float2 s_part[2][4][EPI == 0 ? BLOCK_N : 1];   // per accumulator, per epilogue warp: column sums of its 32 rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = taps * cin_blocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = tile * 128;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                    const int tap = kb / cin_blocks, cb = kb - tap * cin_blocks;
                    const int kh = tap / kwn, kw = tap - kh * kwn;
                    uint8_t* a = tiles + (size_t)s * Cfg::STAGE_BYTES;
                    tma_load_2d(a, &tmA, &full_bar[s], cb * 64, m0 + kh * wp + kw);
                    tma_load_2d(a + Cfg::A_BYTES, &tmB, &full_bar[s], kb * 64, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---- MMA issuer
            // instruction descriptor: D fp32, A/B bf16, both K-major, N = BLOCK_N, M = 128
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t it = 0, t = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                const uint32_t as = t & 1u, aph = (t >> 1) & 1u;
                mbar_wait(&acc_empty[as], aph ^ 1u);       // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * (uint32_t)BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a = smem_u32(tiles + (size_t)s * Cfg::STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(a), bdesc = umma_desc_sw128(a + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)            // 4 x K=16 inside the 128-byte swizzle row: +32 B per step
                        tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                    tc_commit(&empty_bar[s]);              // frees the stage once these MMAs have read it
                }
                tc_commit(&acc_full[as]);                  // accumulator complete
            }
        }
    } else {                                               // ---- epilogue: TMEM -> registers -> global
        const int quarter = warp & 3;                      // TMEM lanes this warp may read
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
            const uint32_t as = t & 1u, aph = (t >> 1) & 1u;
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const int q = tile * 128 + quarter * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * (uint32_t)BLOCK_N;
            const bool stats = EPI == 0 && st.part != nullptr;
            bool mine0 = false, mine1 = false, cross = false;      // this row counts for the tile's first / second image
            int img0 = 0;
            if (stats) {
                const int m0 = tile * 128;
                img0 = m0 / st.hpwp;
                const int r = q - img0 * st.hpwp;
                const bool second = r >= st.hpwp;
                const int rr = second ? r - st.hpwp : r;
                const int hh = rr / wp, ww = rr - hh * wp;
                const bool valid = q < m_total && hh < st.H && ww < st.W;
                mine0 = valid && !second;
                mine1 = valid && second;
                cross = m0 + 127 - img0 * st.hpwp >= st.hpwp;       // (uniform over the tile)
            }
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                if (EPI == 1 && c0 >= cout) break;
                uint32_t v[32];
                tc_ld32(taddr + (uint32_t)c0, v);
                if (EPI == 0) {
                    if (q < m_total) {
                        uint4* d4 = reinterpret_cast<uint4*>(static_cast<bf16*>(out) + (size_t)q * cout + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            d4[j] = make_uint4(pack_bf16(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                               pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                                               pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                               pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
                    }
                    if (stats) {
                        float a[32], b[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) { const float x = mine0 ? __uint_as_float(v[j]) : 0.f; a[j] = x; b[j] = x * x; }
                        const float s0 = warp_colsum(a, lane), s1 = warp_colsum(b, lane);
                        s_part[as][quarter][c0 + lane] = make_float2(s0, s1);
                        if (cross) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) { const float x = mine1 ? __uint_as_float(v[j]) : 0.f; a[j] = x; b[j] = x * x; }
                            const float c0s = warp_colsum(a, lane), c1s = warp_colsum(b, lane);
                            st.carry[((size_t)(img0 + 1) * 4 + quarter) * cout + c0 + lane] = make_float2(c0s, c1s);
                        }
                    }
                } else {
                    if (q < m_total) {
                        float* o = static_cast<float*>(out) + (size_t)c0 * m_total + q;   // a warp writes 32 consecutive q per plane
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + j < cout) o[(size_t)j * m_total] = __uint_as_float(v[j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);    // 4 arrivals (one per epilogue warp) release the accumulator
            if (EPI == 0 && st.part != nullptr) {
                // the four warps' column sums -> one partial per tile, added in a fixed order (s_part is double buffered by
                // accumulator, so the next tile's sums never overwrite what a slower warp still reads)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int col = quarter * 32 + lane; col < BLOCK_N; col += 128) {
                    const float2 p0 = s_part[as][0][col], p1 = s_part[as][1][col], p2 = s_part[as][2][col], p3 = s_part[as][3][col];
                    st.part[(size_t)tile * cout + col] = make_float2((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// input transforms
// ------------------------------------------------------------------------------------------
// per image: min / max of the u8 raster and of the u8 background
__global__ void __launch_bounds__(256) k_gan_minmax(const uint8_t* __restrict__ raster, const uint8_t* __restrict__ bg, int hw, int* __restrict__ mm) {
    const int b = blockIdx.x;
    int mn0 = 255, mx0 = 0, mn1 = 255, mx1 = 0;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
        const int v = raster[(size_t)b * hw + i];
        mn0 = min(mn0, v); mx0 = max(mx0, v);
        if (bg) { const int u = bg[(size_t)b * hw + i]; mn1 = min(mn1, u); mx1 = max(mx1, u); }
    }
    __shared__ int s[4][8];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn0 = min(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mx0 = max(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
        mn1 = min(mn1, __shfl_xor_sync(0xffffffffu, mn1, o)); mx1 = max(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
    }
    if ((threadIdx.x & 31) == 0) { const int w = threadIdx.x >> 5; s[0][w] = mn0; s[1][w] = mx0; s[2][w] = mn1; s[3][w] = mx1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mn0 = min(mn0, s[0][w]); mx0 = max(mx0, s[1][w]); mn1 = min(mn1, s[2][w]); mx1 = max(mx1, s[3][w]); }
        mm[4 * b] = mn0; mm[4 * b + 1] = mx0; mm[4 * b + 2] = mn1; mm[4 * b + 3] = mx1;
    }
}

// legacy numpy stream: np.random.seed(seed); np.random.uniform(0, 1, (H, W)) -- one CTA per image regenerates the
// twister block by block (three dependent phases per 624 words) and tempers two words into one double
__global__ void __launch_bounds__(640) k_gan_speckle(const uint32_t* __restrict__ seeds, int hw, double* __restrict__ out) {
    __shared__ uint32_t st[2][624];
    const int b = blockIdx.x, t = threadIdx.x;
    if (t == 0) {
        uint32_t x = seeds[b];
        st[0][0] = x;
        for (int i = 1; i < 624; ++i) { x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i; st[0][i] = x; }
    }
    __syncthreads();
    const int n_words = 2 * hw;
    int cur = 0;
    for (int base = 0; base < n_words; base += 624) {
        const uint32_t* o = st[cur];
        uint32_t* nw = st[cur ^ 1];
        auto tw = [](uint32_t a, uint32_t bq, uint32_t m) { const uint32_t y = (a & 0x80000000u) | (bq & 0x7fffffffu); return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); };
        if (t < 227) nw[t] = tw(o[t], o[t + 1], o[t + 397]);
        __syncthreads();
        if (t >= 227 && t < 454) nw[t] = tw(o[t], o[t + 1], nw[t - 227]);
        __syncthreads();
        if (t >= 454 && t < 623) nw[t] = tw(o[t], o[t + 1], nw[t - 227]);
        if (t == 623) nw[623] = tw(o[623], nw[0], nw[396]);
        __syncthreads();
        if (t < 312) {
            const int d = base / 2 + t;
            if (d < hw) out[(size_t)b * hw + d] = octa::mt_double(octa::mt_temper(nw[2 * t]), octa::mt_temper(nw[2 * t + 1]));
        }
        cur ^= 1;
        __syncthreads();
    }
}

// x = float32(max(float64(scale(img)), float64(scale(bg^T)) * speckle))      data_transforms.py:507-511
__global__ void k_gan_input(const uint8_t* __restrict__ raster, const uint8_t* __restrict__ bg, const double* __restrict__ speckle,
                            const int* __restrict__ mm, int n, int H, int W, float* __restrict__ x) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t hw = (size_t)H * W;
    if (i >= (size_t)n * hw) return;
    const int b = (int)(i / hw);
    const int p = (int)(i - (size_t)b * hw), h = p / W, w = p - h * W;
    const int mn0 = mm[4 * b], mx0 = mm[4 * b + 1], mn1 = mm[4 * b + 2], mx1 = mm[4 * b + 3];
    const float fi = mx0 > mn0 ? __fdiv_rn((float)raster[i] - (float)mn0, (float)mx0 - (float)mn0) : 0.0f;
    double v = (double)fi;
    if (bg && speckle) {
        const int u = bg[(size_t)b * hw + (size_t)w * W + h];     // Rotate90d(k=1) then Flipd(axis 0) == transpose (square images)
        const float fb = mx1 > mn1 ? __fdiv_rn((float)u - (float)mn1, (float)mx1 - (float)mn1) : 0.0f;
        v = fmax(v, __dmul_rn((double)fb, speckle[i]));
    }
    x[i] = (float)v;
}

// ------------------------------------------------------------------------------------------
// im2col of the 7x7 stem (1 -> 64, reflect pad 3) and the tap sum of the head (64 -> 1 over the materialised reflect pad 3, bias, sigmoid)
// ------------------------------------------------------------------------------------------
// im2col of the single-channel input for the stem: row q of the padded flat grid gets its 49 reflect-padded taps (+ 15 zeros)
// as one 128-byte bf16 row, so the 7x7 stem is ONE GEMM [pixels x 64] x [64 x 64] on the tensor cores (k_gan_conv3, taps = 1)
__global__ void __launch_bounds__(256) k_gan_stem_im2col(const float* __restrict__ x, int n, int H, int W, bf16* __restrict__ a /*[n][(H+2)][(W+2)][64]*/) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int Hp = H + 2, Wp = W + 2;
    if (i >= (size_t)n * Hp * Wp * 8) return;
    const int sg = (int)(i & 7);
    size_t px = i >> 3;
    const int w = (int)(px % Wp); px /= Wp;
    const int h = (int)(px % Hp);
    const int b = (int)(px / Hp);
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h < H && w < W) {
        const float* xb = x + (size_t)b * H * W;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int t = sg * 8 + k;
            if (t < 49) { const int ta = t / 7, tc = t - ta * 7; f[k] = __ldg(xb + (size_t)reflect_idx(h + ta - 3, H) * W + reflect_idx(w + tc - 3, W)); }
        }
    }
    reinterpret_cast<uint4*>(a)[i] = pack8(f);
}

// out(h,w) = sigmoid(bias + sum_{a,b} planes[a*7+b][(h+a, w+b)]) over the [H+6][W+6] grid of the padded head input
__global__ void __launch_bounds__(256) k_gan_head_sum(const float* __restrict__ planes, size_t plane_stride, float bias, int n, int H, int W,
                                                      float* __restrict__ y, uint8_t* __restrict__ y8) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t hw = (size_t)H * W;
    if (i >= (size_t)n * hw) return;
    const int b = (int)(i / hw);
    const int p = (int)(i - (size_t)b * hw), h = p / W, w = p - h * W;
    const int Wp = W + 6;
    const float* base = planes + ((size_t)(b * (H + 6) + h) * Wp + w);
    float acc[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < 7; ++a)
#pragma unroll
        for (int c = 0; c < 7; ++c) acc[a] += __ldg(base + (size_t)(a * 7 + c) * plane_stride + (size_t)a * Wp + c);
    const float z = (((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + acc[6])) + bias;
    const float s = 1.0f / (1.0f + expf(-z));
    if (y) y[i] = s;
    if (y8) y8[i] = (uint8_t)(s * 255.0f);                  // utils/visualizer.py:338 (float32 multiply, truncation)
}

// ------------------------------------------------------------------------------------------
// instance-norm statistics: finalisation of the per-tile partials the conv epilogue leaves (ConvStats)
// ------------------------------------------------------------------------------------------
// mean / rstd per (image, channel) from the per-tile partials of the conv epilogue: tiles whose first row lies in image b, in
// order, plus the four warp partials of the tile that straddles the previous image's end.  fp64 accumulation, biased variance.
__global__ void __launch_bounds__(256) k_gan_stats_tiles(const float2* __restrict__ part, const float2* __restrict__ carry, int C, int hpwp, int hw,
                                                         float2* __restrict__ mr /*[n][C] mean, rstd*/) {
    const int b = blockIdx.x;
    const long long r0 = (long long)b * hpwp, r1 = r0 + hpwp;
    const int t0 = (int)((r0 + 127) / 128), t1 = (int)((r1 + 127) / 128);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double s = 0, ss = 0;
        if (r0 % 128 != 0)
            for (int w = 0; w < 4; ++w) { const float2 p = carry[((size_t)b * 4 + w) * C + c]; s += p.x; ss += p.y; }
#pragma unroll 8
        for (int t = t0; t < t1; ++t) { const float2 p = __ldg(part + (size_t)t * C + c); s += p.x; ss += p.y; }
        const double mean = s / hw;
        double var = ss / hw - mean * mean;
        if (var < 0) var = 0;
        mr[(size_t)b * C + c] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
    }
}

// raw (flat rows of the [H+2][W+2] grid) -> (x - mean) * rstd [ReLU] [+ skip]  -> destination buffer with padding P
__global__ void __launch_bounds__(256) k_gan_norm(const bf16* __restrict__ raw, const float2* __restrict__ mr, const bf16* __restrict__ skip,
                                                  int n, int H, int W, int C, int P, int reflect, int relu, bf16* __restrict__ dst) {
    const int groups = C >> 3;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int Hd = H + 2 * P, Wd = W + 2 * P;
    if (i >= (size_t)n * Hd * Wd * groups) return;
    const int cg = (int)(i % groups);
    size_t px = i / groups;
    const int wp = (int)(px % Wd); px /= Wd;
    const int hp = (int)(px % Hd);
    const int b = (int)(px / Hd);
    int h = hp - P, w = wp - P;
    uint4 out = make_uint4(0, 0, 0, 0);
    const bool inside = h >= 0 && h < H && w >= 0 && w < W;
    if (inside || reflect) {
        h = reflect_idx(h, H); w = reflect_idx(w, W);
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(raw + ((size_t)(b * (H + 2) + h) * (W + 2) + w) * C) + cg), f);
        const float4* m4 = reinterpret_cast<const float4*>(mr + (size_t)b * C + cg * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 m = __ldg(m4 + k);
            f[2 * k] = (f[2 * k] - m.x) * m.y;
            f[2 * k + 1] = (f[2 * k + 1] - m.z) * m.w;
        }
        if (relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
        }
        if (skip) {
            float g[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(skip + ((size_t)(b * (H + 2) + h + 1) * (W + 2) + w + 1) * C) + cg), g);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += g[k];
        }
        out = pack8(f);
    }
    reinterpret_cast<uint4*>(dst)[i] = out;
}

// anti-aliased resampling into a padded (P = 1) buffer: mode 0 = Downsample (networks.py:266-289: reflect pad 1,
// [1,2,1]^2/16, stride 2), mode 1 = Upsample (networks.py:244-264: replicate pad 1, [1,3,3,1]^2/64*4 transposed, cropped).
// The source is either a padded activation buffer (mr == nullptr) or a RAW conv output whose InstanceNorm + ReLU is applied
// per tap on the fly (mr = mean / rstd per image and channel) -- the normalised full-resolution tensor is never written.
__global__ void __launch_bounds__(256) k_gan_resample(const bf16* __restrict__ src, const float2* __restrict__ mr, int n, int Hs, int Ws, int C,
                                                      int mode, int reflect, bf16* __restrict__ dst) {
    const int groups = C >> 3;
    const int H = mode == 0 ? Hs / 2 : Hs * 2, W = mode == 0 ? Ws / 2 : Ws * 2;
    const int Hd = H + 2, Wd = W + 2;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * Hd * Wd * groups) return;
    const int cg = (int)(i % groups);
    size_t px = i / groups;
    const int wp = (int)(px % Wd); px /= Wd;
    const int hp = (int)(px % Hd);
    const int b = (int)(px / Hd);
    int h = hp - 1, w = wp - 1;
    uint4 out = make_uint4(0, 0, 0, 0);
    const bool inside = h >= 0 && h < H && w >= 0 && w < W;
    if (inside || reflect) {
        h = reflect_idx(h, H); w = reflect_idx(w, W);
        int ih[3], iw[3];
        float fh[3], fw[3];
        int taps;
        if (mode == 0) {
            taps = 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) { ih[k] = reflect_idx(2 * h - 1 + k, Hs); iw[k] = reflect_idx(2 * w - 1 + k, Ws); }
            fh[0] = fw[0] = 0.25f; fh[1] = fw[1] = 0.5f; fh[2] = fw[2] = 0.25f;
        } else {
            taps = 2;
            if (h & 1) { ih[0] = (h - 1) / 2; ih[1] = min((h + 1) / 2, Hs - 1); } else { ih[0] = h / 2; ih[1] = max(h / 2 - 1, 0); }
            if (w & 1) { iw[0] = (w - 1) / 2; iw[1] = min((w + 1) / 2, Ws - 1); } else { iw[0] = w / 2; iw[1] = max(w / 2 - 1, 0); }
            fh[0] = fw[0] = 0.75f; fh[1] = fw[1] = 0.25f; fh[2] = fw[2] = 0.f; ih[2] = iw[2] = 0;
        }
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float mean[8], rstd[8];
        const int off = mr ? 0 : 1;                          // raw outputs sit at (h, w) of the grid, activations at (h+1, w+1)
        if (mr) {
            const float4* m4 = reinterpret_cast<const float4*>(mr + (size_t)b * C + cg * 8);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float4 m = __ldg(m4 + k); mean[2 * k] = m.x; rstd[2 * k] = m.y; mean[2 * k + 1] = m.z; rstd[2 * k + 1] = m.w; }
        }
        for (int a = 0; a < taps; ++a)
            for (int c = 0; c < taps; ++c) {
                float f[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(src + ((size_t)(b * (Hs + 2) + ih[a] + off) * (Ws + 2) + iw[c] + off) * C) + cg), f);
                if (mr) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] = fmaxf((f[k] - mean[k]) * rstd[k], 0.f);
                }
                const float wgt = fh[a] * fw[c];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, f[k], acc[k]);
            }
        out = pack8(acc);
    }
    reinterpret_cast<uint4*>(dst)[i] = out;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// 2-D bf16 tensor [rows][inner], box [box_rows][64], 128-byte swizzle, zero fill outside
int make_map(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { octa::set_error("cuTensorMapEncodeTiled is not available from this driver"); return OCTA_E_CUDA; }
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {inner * sizeof(bf16)};
    const cuuint32_t box[2] = {64, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { octa::set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return OCTA_E_CUDA; }
    return OCTA_OK;
}

struct Conv3 {            // one 3x3 layer: weights [cout][9][cin] bf16 and its TMA maps
    int cin = 0, cout = 0;
    bf16* w = nullptr;
    CUtensorMap tmB;
};

struct GanCtx {
    int max_n = 0, H = 0, W = 0, n_sm = 148, two_ctas = -1;
    float* planes = nullptr;
    Conv3 stem;           // [64 cout][64 slots: tap t < 49, zero above] bf16 -- the stem as one GEMM over im2col rows
    Conv3 head;           // [64 rows: tap t < 49, zero above][64 channels] bf16 -- the head as one GEMM
    float head_b = 0.f;
    Conv3 conv[22];       // 0,1 = down; 2..19 = blocks (a, b); 20,21 = up
    // activations (bf16, padded), see octa_gan_forward_dev
    bf16 *raw = nullptr, *a1 = nullptr, *a2 = nullptr, *r0 = nullptr, *r1 = nullptr, *rt = nullptr, *u1 = nullptr, *b3 = nullptr, *u2 = nullptr, *fin = nullptr;
    float2 *mr = nullptr, *part = nullptr, *carry = nullptr;
    std::vector<void*> allocs;
    ~GanCtx() { for (void* p : allocs) cudaFree(p); }
};

template <class T> int dev_alloc(GanCtx* c, T** p, size_t count) {
    void* q = nullptr;
    if (cudaMalloc(&q, count * sizeof(T)) != cudaSuccess) { cudaGetLastError(); octa::set_error("cudaMalloc of %zu bytes failed", count * sizeof(T)); return OCTA_E_NOMEM; }
    c->allocs.push_back(q);
    *p = (T*)q;
    return OCTA_OK;
}

uint16_t f2bf(float f) {   // round to nearest even
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

int upload_conv(GanCtx* c, Conv3* L, const float* w /*[cout][cin][3][3]*/, int cin, int cout) {
    L->cin = cin; L->cout = cout;
    std::vector<uint16_t> h((size_t)cout * 9 * cin);
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int t = 0; t < 9; ++t) h[((size_t)co * 9 + t) * cin + ci] = f2bf(w[((size_t)co * cin + ci) * 9 + t]);
    int rc = dev_alloc(c, &L->w, h.size());
    if (rc) return rc;
    OCTA_CUDA_CHECK(cudaMemcpy(L->w, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    return make_map(&L->tmB, L->w, (uint64_t)9 * cin, (uint64_t)cout, (uint32_t)cout);
}

int upload_stem(GanCtx* c, const float* w /*[64][1][7][7]*/) {
    c->stem.cin = 64; c->stem.cout = 64;
    std::vector<uint16_t> h(64 * 64, 0);
    for (int co = 0; co < 64; ++co)
        for (int t = 0; t < 49; ++t) h[co * 64 + t] = f2bf(w[co * 49 + t]);
    int rc = dev_alloc(c, &c->stem.w, h.size());
    if (rc) return rc;
    OCTA_CUDA_CHECK(cudaMemcpy(c->stem.w, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    return make_map(&c->stem.tmB, c->stem.w, 64, 64, 64);
}

int upload_head(GanCtx* c, const float* w /*[1][64][7][7]*/) {
    c->head.cin = 64; c->head.cout = 49;
    std::vector<uint16_t> h(64 * 64, 0);
    for (int t = 0; t < 49; ++t)
        for (int ch = 0; ch < 64; ++ch) h[t * 64 + ch] = f2bf(w[ch * 49 + t]);
    int rc = dev_alloc(c, &c->head.w, h.size());
    if (rc) return rc;
    OCTA_CUDA_CHECK(cudaMemcpy(c->head.w, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    return make_map(&c->head.tmB, c->head.w, 64, 64, 64);
}


// raw = conv3x3(act) over the flat rows of act's padded grid
int run_conv(const GanCtx* c, const Conv3& L, const bf16* act, int n, int H, int W, cudaStream_t st, bool with_stats = true) {
    const int Wp = W + 2;
    const long long rows = (long long)n * (H + 2) * Wp;
    CUtensorMap tmA;
    int rc = make_map(&tmA, act, (uint64_t)L.cin, (uint64_t)rows, 128);
    if (rc) return rc;
    const int mt = (int)((rows + 127) / 128);
    const ConvStats cs = {with_stats ? c->part : nullptr, c->carry, (H + 2) * Wp, H, W};
    auto grid = [&](int ctas) { const int g = c->n_sm * ctas; return (unsigned)(mt < g ? mt : g); };
    // bit 0: the Cout = 128 layers, bit 1: the Cout = 64 layer run 2 CTAs per SM (half the ring each).  Their A tiles feed fewer
    // MMA columns, so they are bound by L2 -> shared-memory traffic and two producers keep more of it in flight (measured per
    // layer on B200, 1 vs 2 CTAs: 551 -> 504 us, 560 -> 385 us, 1008 -> 577 us); the Cout = 256 layers prefer one CTA with N = 256
    const int two = c->two_ctas >= 0 ? c->two_ctas : 3;
    if (L.cout == 256)
        k_gan_conv3<256, 0, 1><<<grid(1), CONV_THREADS, ConvCfg<256, 1>::SMEM, st>>>(tmA, L.tmB, c->raw, (int)rows, Wp, L.cin / 64, L.cout, 9, 3, mt, cs);
    else if (L.cout == 128 && !(two & 1))
        k_gan_conv3<128, 0, 1><<<grid(1), CONV_THREADS, ConvCfg<128, 1>::SMEM, st>>>(tmA, L.tmB, c->raw, (int)rows, Wp, L.cin / 64, L.cout, 9, 3, mt, cs);
    else if (L.cout == 128)
        k_gan_conv3<128, 0, 2><<<grid(2), CONV_THREADS, ConvCfg<128, 2>::SMEM, st>>>(tmA, L.tmB, c->raw, (int)rows, Wp, L.cin / 64, L.cout, 9, 3, mt, cs);
    else if (!(two & 2))
        k_gan_conv3<64, 0, 1><<<grid(1), CONV_THREADS, ConvCfg<64, 1>::SMEM, st>>>(tmA, L.tmB, c->raw, (int)rows, Wp, L.cin / 64, L.cout, 9, 3, mt, cs);
    else
        k_gan_conv3<64, 0, 2><<<grid(2), CONV_THREADS, ConvCfg<64, 2>::SMEM, st>>>(tmA, L.tmB, c->raw, (int)rows, Wp, L.cin / 64, L.cout, 9, 3, mt, cs);
    octa::count_launch();
    if (with_stats) {
        k_gan_stats_tiles<<<n, 256, 0, st>>>(c->part, c->carry, L.cout, (H + 2) * Wp, H * W, c->mr);
        octa::count_launch();
    }
    OCTA_CUDA_CHECK(cudaGetLastError());
    return OCTA_OK;
}

int run_norm(const GanCtx* c, const bf16* skip, int n, int H, int W, int C, int P, int reflect, int relu, bf16* dst, cudaStream_t st) {
    const size_t total = (size_t)n * (H + 2 * P) * (W + 2 * P) * (C / 8);
    k_gan_norm<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(c->raw, c->mr, skip, n, H, W, C, P, reflect, relu, dst);
    octa::count_launch();
    OCTA_CUDA_CHECK(cudaGetLastError());
    return OCTA_OK;
}

int run_resample(const bf16* src, const float2* mr, int n, int Hs, int Ws, int C, int mode, int reflect, bf16* dst, cudaStream_t st) {
    const int H = mode == 0 ? Hs / 2 : Hs * 2, W = mode == 0 ? Ws / 2 : Ws * 2;
    const size_t total = (size_t)n * (H + 2) * (W + 2) * (C / 8);
    k_gan_resample<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, mr, n, Hs, Ws, C, mode, reflect, dst);
    octa::count_launch();
    OCTA_CUDA_CHECK(cudaGetLastError());
    return OCTA_OK;
}

int conv_attrs(GanCtx* c) {
    int dev = 0, sm = 0;
    OCTA_CUDA_CHECK(cudaGetDevice(&dev));
    OCTA_CUDA_CHECK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    c->n_sm = sm > 0 ? sm : 148;
    const char* e = getenv("OCTA_GAN_TWO_CTAS");             // diagnostics: which layers run 2 CTAs per SM (see run_conv)
    if (e) c->two_ctas = atoi(e);
    OCTA_CUDA_CHECK(cudaFuncSetAttribute(k_gan_conv3<256, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvCfg<256, 1>::SMEM));
    OCTA_CUDA_CHECK(cudaFuncSetAttribute(k_gan_conv3<128, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvCfg<128, 1>::SMEM));
    OCTA_CUDA_CHECK(cudaFuncSetAttribute(k_gan_conv3<128, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvCfg<128, 2>::SMEM));
    OCTA_CUDA_CHECK(cudaFuncSetAttribute(k_gan_conv3<64, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvCfg<64, 1>::SMEM));
    OCTA_CUDA_CHECK(cudaFuncSetAttribute(k_gan_conv3<64, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvCfg<64, 2>::SMEM));
    OCTA_CUDA_CHECK(cudaFuncSetAttribute(k_gan_conv3<64, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvCfg<64, 1>::SMEM));
    return OCTA_OK;
}

#define GAN_TRY(expr) do { const int _rc = (expr); if (_rc) return _rc; } while (0)

}  // namespace

extern "C" int octa_gan_create(const OctaGanWeights* w, int max_images, int H, int W, void** handle) {
    OCTA_ARG_CHECK(w && handle, "null argument");
    OCTA_ARG_CHECK(max_images > 0 && max_images <= 4096, "max_images out of range");
    OCTA_ARG_CHECK(H >= 16 && W >= 16 && H % 4 == 0 && W % 4 == 0 && H <= 4096 && W <= 4096, "H and W must be multiples of 4 in [16, 4096]");
    OCTA_ARG_CHECK((long long)max_images * (H + 6) * (W + 6) < (1ll << 31) - 4096, "max_images * (H+6) * (W+6) must stay below 2^31");
    OCTA_ARG_CHECK(w->stem_w && w->head_w, "null weight pointer");
    for (int i = 0; i < 22; ++i) OCTA_ARG_CHECK(w->conv_w[i], "null weight pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); octa::set_error("no CUDA device"); return OCTA_E_CUDA; }
    GanCtx* c = new (std::nothrow) GanCtx();
    if (!c) { octa::set_error("out of host memory"); return OCTA_E_NOMEM; }
    c->max_n = max_images; c->H = H; c->W = W; c->head_b = w->head_b;
    int rc = OCTA_OK;
    auto fail = [&](int code) { delete c; return code; };
    if ((rc = upload_stem(c, w->stem_w))) return fail(rc);
    if ((rc = upload_head(c, w->head_w))) return fail(rc);
    static const int cio[22][2] = {{64, 128}, {128, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256},
                                   {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256}, {256, 256},
                                   {256, 128}, {128, 64}};
    for (int i = 0; i < 22; ++i)
        if ((rc = upload_conv(c, &c->conv[i], w->conv_w[i], cio[i][0], cio[i][1]))) return fail(rc);
    const size_t n = (size_t)max_images;
    const size_t p1 = (size_t)(H + 2) * (W + 2), p2 = (size_t)(H / 2 + 2) * (W / 2 + 2), p3 = (size_t)(H / 4 + 2) * (W / 4 + 2);
    const size_t slack = 128 * 256;       // (TMA boxes past the last row are zero-filled; the slack only keeps stores of partial tiles simple)
    if ((rc = dev_alloc(c, &c->raw, n * p1 * 128 + slack)) || (rc = dev_alloc(c, &c->a1, n * p1 * 64)) ||
        (rc = dev_alloc(c, &c->a2, n * p2 * 128)) || (rc = dev_alloc(c, &c->r0, n * p3 * 256)) ||
        (rc = dev_alloc(c, &c->r1, n * p3 * 256)) || (rc = dev_alloc(c, &c->rt, n * p3 * 256)) || (rc = dev_alloc(c, &c->u1, n * p2 * 256)) ||
        (rc = dev_alloc(c, &c->b3, n * p2 * 128)) || (rc = dev_alloc(c, &c->u2, n * p1 * 128)) ||
        (rc = dev_alloc(c, &c->fin, n * (size_t)(H + 6) * (W + 6) * 64)) || (rc = dev_alloc(c, &c->planes, n * (size_t)(H + 6) * (W + 6) * 49)) || (rc = dev_alloc(c, &c->part, (n * p1 / 128 + 2) * 128)) ||
        (rc = dev_alloc(c, &c->carry, (n + 1) * 4 * 256)) || (rc = dev_alloc(c, &c->mr, n * 256)))
        return fail(rc);
    if ((rc = conv_attrs(c))) return fail(rc);
    *handle = c;
    return OCTA_OK;
}

extern "C" void octa_gan_destroy(void* handle) { delete static_cast<GanCtx*>(handle); }

extern "C" int octa_gan_forward_dev(void* handle, const float* x_dev, int n_images, float* y_dev, uint8_t* y_u8_dev, void* stream) {
    GanCtx* c = static_cast<GanCtx*>(handle);
    OCTA_ARG_CHECK(c && x_dev && (y_dev || y_u8_dev), "null argument");
    OCTA_ARG_CHECK(n_images > 0 && n_images <= c->max_n, "n_images exceeds the context's max_images");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = n_images, H = c->H, W = c->W, H2 = H / 2, W2 = W / 2, H3 = H / 4, W3 = W / 4;
    // stem: ReflectionPad2d(3) + Conv2d(1, 64, 7) + IN + ReLU                                 networks.py:372-375
    // im2col rows (49 taps of the one input channel, bf16) in u2's storage, then one tensor-core GEMM into raw
    {
        const long long rows = (long long)n * (H + 2) * (W + 2);
        k_gan_stem_im2col<<<(unsigned)((rows * 8 + 255) / 256), 256, 0, st>>>(x_dev, n, H, W, c->u2);
        CUtensorMap tmA;
        GAN_TRY(make_map(&tmA, c->u2, 64, (uint64_t)rows, 128));
        const int mt = (int)((rows + 127) / 128);
        const int g = 2 * c->n_sm;
        k_gan_conv3<64, 0, 2><<<(unsigned)(mt < g ? mt : g), CONV_THREADS, ConvCfg<64, 2>::SMEM, st>>>(tmA, c->stem.tmB, c->raw, (int)rows, W + 2, 1, 64, 1, 1, mt,
                                                                                                              ConvStats{c->part, c->carry, (H + 2) * (W + 2), H, W});
        k_gan_stats_tiles<<<n, 256, 0, st>>>(c->part, c->carry, 64, (H + 2) * (W + 2), H * W, c->mr);
        octa::count_launch(3);
        OCTA_CUDA_CHECK(cudaGetLastError());
    }
    GAN_TRY(run_norm(c, nullptr, n, H, W, 64, 1, 0, 1, c->a1, st));
    // down 1: Conv2d(64, 128, 3, padding=1) + IN + ReLU + Downsample                          networks.py:384-387
    GAN_TRY(run_conv(c, c->conv[0], c->a1, n, H, W, st));
    GAN_TRY(run_resample(c->raw, c->mr, n, H, W, 128, 0, 0, c->a2, st));        // IN + ReLU applied per tap
    // down 2
    GAN_TRY(run_conv(c, c->conv[1], c->a2, n, H2, W2, st));
    GAN_TRY(run_resample(c->raw, c->mr, n, H2, W2, 256, 0, 1, c->r0, st));
    // 9 x ResnetBlock: x + IN(conv(pad(ReLU(IN(conv(pad(x)))))))                               networks.py:291-348
    bf16 *cur = c->r0, *nxt = c->r1;
    for (int blk = 0; blk < 9; ++blk) {
        GAN_TRY(run_conv(c, c->conv[2 + 2 * blk], cur, n, H3, W3, st));
        GAN_TRY(run_norm(c, nullptr, n, H3, W3, 256, 1, 1, 1, c->rt, st));
        GAN_TRY(run_conv(c, c->conv[3 + 2 * blk], c->rt, n, H3, W3, st));
        GAN_TRY(run_norm(c, cur, n, H3, W3, 256, 1, 1, 0, nxt, st));
        bf16* t = cur; cur = nxt; nxt = t;
    }
    // up 1 / up 2: Upsample + Conv2d(3, padding=1) + IN + ReLU                                networks.py:408-414
    GAN_TRY(run_resample(cur, nullptr, n, H3, W3, 256, 1, 0, c->u1, st));
    GAN_TRY(run_conv(c, c->conv[20], c->u1, n, H2, W2, st));
    // (up-sampling makes 4 outputs per input: normalising once into b3 is cheaper than per tap -- 650 vs 840 us per 32 images)
    GAN_TRY(run_norm(c, nullptr, n, H2, W2, 128, 1, 0, 1, c->b3, st));
    GAN_TRY(run_resample(c->b3, nullptr, n, H2, W2, 128, 1, 0, c->u2, st));
    GAN_TRY(run_conv(c, c->conv[21], c->u2, n, H, W, st));
    GAN_TRY(run_norm(c, nullptr, n, H, W, 64, 3, 1, 1, c->fin, st));
    // head: ReflectionPad2d(3) + Conv2d(64, 1, 7) + Sigmoid                                    networks.py:415-417
    // one GEMM [pixels of the padded grid x 64 ch] x [64 ch x 49 taps] on the tensor cores, then the shifted sum of the 49 planes
    {
        const long long rows = (long long)n * (H + 6) * (W + 6);
        CUtensorMap tmA;
        GAN_TRY(make_map(&tmA, c->fin, 64, (uint64_t)rows, 128));
        const int mt = (int)((rows + 127) / 128);
        k_gan_conv3<64, 1, 1><<<(unsigned)(mt < c->n_sm ? mt : c->n_sm), CONV_THREADS, ConvCfg<64, 1>::SMEM, st>>>(tmA, c->head.tmB, c->planes, (int)rows, W + 6, 1, 49, 1, 1, mt, ConvStats{nullptr, nullptr, 1, 0, 0});
        const size_t total = (size_t)n * H * W;
        k_gan_head_sum<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(c->planes, (size_t)rows, c->head_b, n, H, W, y_dev, y_u8_dev);
        octa::count_launch(2);
        OCTA_CUDA_CHECK(cudaGetLastError());
    }
    return OCTA_OK;
}

extern "C" int octa_gan_speckle_dev(const uint32_t* seeds_dev, int n_images, int H, int W, double* speckle_dev, void* stream) {
    OCTA_ARG_CHECK(seeds_dev && speckle_dev && n_images > 0 && H > 0 && W > 0, "bad argument");
    k_gan_speckle<<<n_images, 640, 0, (cudaStream_t)stream>>>(seeds_dev, H * W, speckle_dev);
    octa::count_launch();
    OCTA_CUDA_CHECK(cudaGetLastError());
    return OCTA_OK;
}

extern "C" int octa_gan_input_dev(const uint8_t* raster_dev, const uint8_t* background_dev, const double* speckle_dev, int n_images, int H, int W,
                                  int* minmax_ws_dev, float* x_dev, void* stream) {
    OCTA_ARG_CHECK(raster_dev && x_dev && minmax_ws_dev && n_images > 0 && H > 0 && W > 0, "bad argument");
    OCTA_ARG_CHECK(!background_dev || H == W, "the background transpose needs square images");
    OCTA_ARG_CHECK((background_dev == nullptr) == (speckle_dev == nullptr), "background and speckle go together");
    cudaStream_t st = (cudaStream_t)stream;
    k_gan_minmax<<<n_images, 256, 0, st>>>(raster_dev, background_dev, H * W, minmax_ws_dev);
    const size_t total = (size_t)n_images * H * W;
    k_gan_input<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(raster_dev, background_dev, speckle_dev, minmax_ws_dev, n_images, H, W, x_dev);
    octa::count_launch(2);
    OCTA_CUDA_CHECK(cudaGetLastError());
    return OCTA_OK;
}

// Test hook: one 3x3 convolution through k_gan_conv3 with HOST tensors in the reference's layouts
// (x [n][Cin][H][W], w [Cout][Cin][3][3], y [n][Cout][H][W]; zero or reflect padding 1).  Used by tests/test_gan_gpu.py to
// check the tensor-core kernel alone against torch.nn.functional.conv2d.
extern "C" int octa_test_gan_conv3_host(const float* x, const float* w, int n, int H, int W, int cin, int cout, int reflect, float* y) {
    OCTA_ARG_CHECK(x && w && y && n > 0 && H > 1 && W > 1, "bad argument");
    OCTA_ARG_CHECK(cin % 64 == 0 && (cout == 64 || cout == 128 || cout == 256), "Cin must be a multiple of 64, Cout 64, 128 or 256");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); octa::set_error("no CUDA device"); return OCTA_E_CUDA; }
    GanCtx c;
    Conv3 L;
    int rc = upload_conv(&c, &L, w, cin, cout);
    if (rc) return rc;
    const int Hp = H + 2, Wp = W + 2;
    const size_t rows = (size_t)n * Hp * Wp;
    std::vector<uint16_t> act(rows * cin);
    for (int b = 0; b < n; ++b)
        for (int hp = 0; hp < Hp; ++hp)
            for (int wp = 0; wp < Wp; ++wp) {
                int h = hp - 1, ww = wp - 1;
                const bool inside = h >= 0 && h < H && ww >= 0 && ww < W;
                if (reflect) { h = h < 0 ? -h : (h >= H ? 2 * H - 2 - h : h); ww = ww < 0 ? -ww : (ww >= W ? 2 * W - 2 - ww : ww); }
                for (int ci = 0; ci < cin; ++ci)
                    act[(((size_t)b * Hp + hp) * Wp + wp) * cin + ci] =
                        (inside || reflect) ? f2bf(x[(((size_t)b * cin + ci) * H + h) * W + ww]) : (uint16_t)0;
            }
    bf16* d_act = nullptr;
    if ((rc = dev_alloc(&c, &d_act, act.size())) || (rc = dev_alloc(&c, &c.raw, rows * cout + 128 * 256))) return rc;
    OCTA_CUDA_CHECK(cudaMemcpy(d_act, act.data(), act.size() * 2, cudaMemcpyHostToDevice));
    if ((rc = conv_attrs(&c))) return rc;
    if ((rc = run_conv(&c, L, d_act, n, H, W, 0, false))) return rc;
    OCTA_CUDA_CHECK(cudaDeviceSynchronize());
    std::vector<uint16_t> raw(rows * cout);
    OCTA_CUDA_CHECK(cudaMemcpy(raw.data(), c.raw, raw.size() * 2, cudaMemcpyDeviceToHost));
    for (int b = 0; b < n; ++b)
        for (int co = 0; co < cout; ++co)
            for (int h = 0; h < H; ++h)
                for (int ww = 0; ww < W; ++ww) {
                    const uint32_t u = (uint32_t)raw[(((size_t)b * Hp + h) * Wp + ww) * cout + co] << 16;
                    float f;
                    memcpy(&f, &u, 4);
                    y[(((size_t)b * cout + co) * H + h) * W + ww] = f;
                }
    return OCTA_OK;
}
