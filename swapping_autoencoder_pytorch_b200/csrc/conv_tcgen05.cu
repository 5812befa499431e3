// Implicit-GEMM convolution on the 5th-generation tensor cores: TMA-fed tcgen05.mma (kind::tf32), fp32
// accumulators in TMEM, fused epilogue, TMA store.  sm_100a only.
//
//   y[n,p,q,k] = sum_{tap t} sum_c  src[n, p*st + oy_t, q*st + ox_t, c] * wmat[k, t*C + c]
//
// GEMM view: M = output pixels, N = output channels, K = taps x channels.  One CTA computes a
// 128 (pixels) x BLOCK_N (channels) tile:
//   * the 128 pixels are a (tn x th x tw) box of the NHWC output; for tap t the matching A tile is the SAME box of
//     the input shifted by (oy_t, ox_t) — fetched with ONE 4-D TMA (box 32ch x tw x th x tn, element stride = conv
//     stride), out-of-bounds rows/columns zero-filled by the TMA unit (that is the conv's zero padding), landing in
//     shared memory directly in the K-major 128-byte-swizzled layout tcgen05 consumes: no im2col buffer, no
//     register staging;
//   * B tile = 32 (k) x BLOCK_N rows of the [Cout, taps*C] filter matrix, 2-D TMA, same swizzle;
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator, warps 2..5 = epilogue
//     (tcgen05.ld 32 lanes x 32 columns per warp -> bias / noise / leaky-ReLU / residual / TF32 rounding -> swizzled
//     staging -> TMA store, which also clips ragged tile edges);
//   * STAGES-deep mbarrier ring (full/empty), tcgen05.commit releases a stage as soon as its MMAs retire.
// fprop uses it with taps (r - pad_t, s - pad_l); stride-1 dgrad with taps (pad_t - r, pad_l - s) over dy and the
// [C, R*S*K] transposed filter.  Operands are consumed at TF32 precision (low 13 mantissa bits ignored by the tensor
// core); producers in this library round-to-nearest to TF32 so that this truncation is exact.
#include "tc_common.cuh"
#include <mutex>

namespace sae {

// ------------------------------------------------------------------------------------------------ host: driver API
EncodeTiledFn g_encode = nullptr;
static bool g_tc_ok = false;

static void tc_init_once() {
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return;
        if (prop.major != 10) return;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess) return;
        if (qres != cudaDriverEntryPointSuccess || fn == nullptr) return;
        g_encode = (EncodeTiledFn)fn;
        const char* off = getenv("SAE_DISABLE_TCGEN05");
        g_tc_ok = !(off && off[0] == '1');
    });
}

bool tc_available() {
    tc_init_once();
    return g_tc_ok;
}

// ------------------------------------------------------------------------------------------------ kernel
constexpr int TC_MAX_TAPS = 16;
constexpr int TC_THREADS = 192;
constexpr int TC_BK = 32;                      // fp32 elements per K block = one 128-byte swizzle row
constexpr int TC_A_BYTES = 128 * TC_BK * 4;    // 16 KB

struct TcParams {
    int num_cblk;                 // source channels / 32
    int ntaps;
    int stride;
    int tw, th, tn;               // output tile box (tw*th*tn == 128)
    int tiles_w, tiles_h, tiles_n;
    int ON, OH, OW, Ncol;         // output [ON, OH, OW, Ncol]
    int src_c;                    // source channels
    int o_mul, o_offy, o_offx;    // output sub-grid -> full-resolution pixel (strided outputs of transposed conv)
    int FH, FW;                   // full-resolution output extent (for noise / residual addressing)
    short oy[TC_MAX_TAPS], ox[TC_MAX_TAPS];   // source pixel = stride * output pixel + (oy, ox)
    int wk[TC_MAX_TAPS];          // K offset of the tap's filter slice inside a wmat row
    // shared-window kernel only: one (wh x ww)-pixel input window per channel block serves every tap
    int ww, wh, oy_min, ox_min;
    unsigned short arow[TC_MAX_TAPS];   // first window row of tap t: (oy_t - oy_min) * ww + (ox_t - ox_min)
    int w_nstride;                      // shared-window kernel: filter rows between consecutive images (0 = one filter for the
                                        // batch; Ncol = per-sample filters [N, Ncol, Ktot], the style-modulated convolution)
    int b_resident;                     // shared-window kernel: the whole filter slice of this CTA's column block (ntaps x
                                        // num_cblk tiles) fits the B ring and is loaded ONCE per CTA instead of once per pixel tile
    int debug;                          // timing experiments (SAE_TC_DEBUG, conv_tc5m only): 1 no output stores, 2 no epilogue work, 4 no MMAs
    EpiParams epi;
};

// ------------------------------------------------------------------------------------------------ epilogue pieces shared by all kernels
// One 32-column chunk of one accumulator row (= one output pixel): bias / NoiseInjection / leaky-ReLU / gain / residual merge /
// TF32 rounding on the 32 values a thread read from TMEM, then the 128-byte row into the chunk's staging tile
// ([128 rows][128 bytes], 16-byte pieces XOR-swizzled by (row & 7): the layout the SWIZZLE_128B output tensor map reads).
__device__ __forceinline__ void tc_epilogue_math(float (&v)[32], const EpiParams& e, int colb, int64_t pixel, int ncol, bool valid, float nz) {
    uint32_t pos = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float t = v[j];
        if (e.bias) t += __ldg(e.bias + colb + j);
        t += nz;
        pos |= (t > 0.f ? 1u : 0u) << j;
        if (e.act == 3) t = t > 0.f ? t : t * e.alpha;
        t *= e.gain;
        v[j] = t;
    }
    // activation bit mask: the thread holds the 32 consecutive channels of one pixel = exactly one word.  The backward passes
    // then read 1 bit instead of 32 per element to learn the leaky-ReLU branch (sae_bias_act_backward, sae_fir_act_backward)
    if (e.act_mask && valid) e.act_mask[(pixel * ncol + colb) >> 5] = pos;
    if (e.residual && valid) {
        const float4* r4 = reinterpret_cast<const float4*>(e.residual + pixel * ncol + colb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 r = __ldg(r4 + j);
            v[4 * j + 0] = (v[4 * j + 0] + r.x) * e.res_scale;
            v[4 * j + 1] = (v[4 * j + 1] + r.y) * e.res_scale;
            v[4 * j + 2] = (v[4 * j + 2] + r.z) * e.res_scale;
            v[4 * j + 3] = (v[4 * j + 3] + r.w) * e.res_scale;
        }
    }
    if (e.round_tf32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rna_tf32(v[j]);
    }
}

__device__ __forceinline__ void tc_stage_row(uint8_t* stg_row, int row, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        *reinterpret_cast<float4*>(stg_row + ((j ^ (row & 7)) << 4)) = o;
    }
}

template <int BLOCK_N>
constexpr int tc_stages() { return BLOCK_N >= 128 ? 3 : 4; }   // <= 96 KB of ring per CTA: two CTAs co-reside on an SM,
                                                                // one runs its epilogue while the other feeds the tensor core

template <int BLOCK_N>
constexpr size_t tc_smem_bytes() {
    // stage ring (A + B per stage); the epilogue staging (BLOCK_N/32 chunks of 16 KB) reuses it after the main loop.
    size_t ring = (size_t)tc_stages<BLOCK_N>() * (TC_A_BYTES + BLOCK_N * 128);
    size_t epi = (size_t)(BLOCK_N / 32) * TC_A_BYTES;
    return (ring > epi ? ring : epi) + 1024 /*alignment slack*/ + 256 /*barriers*/;
}

// Up to TC_MAX_BATCH problems over the same filter matrix in one launch (the thin remainder strips of the four parity classes of
// a stride-2 data gradient: each is a handful of tiles whose cost is the latency of one K loop — launched one after the
// other they cost 4 x 20..30 us behind a 0.5 ms main kernel, together one).  blockIdx.x walks the problems' tiles back to back.
constexpr int TC_MAX_BATCH = 4;
struct TcBatch {
    int count;
    int tile_end[TC_MAX_BATCH];               // running sum of the problems' pixel-tile counts
    TcParams p[TC_MAX_BATCH];
    CUtensorMap src[TC_MAX_BATCH], out[TC_MAX_BATCH];
    CUtensorMap w;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_tc_kernel(const __grid_constant__ TcBatch batch) {
    int prob = 0;
    while (prob + 1 < batch.count && (int)blockIdx.x >= batch.tile_end[prob]) ++prob;
    const TcParams& p = batch.p[prob];
    const CUtensorMap& map_src = batch.src[prob];
    const CUtensorMap& map_out = batch.out[prob];
    const CUtensorMap& map_w = batch.w;
    constexpr int STAGES = tc_stages<BLOCK_N>();
    constexpr int B_BYTES = BLOCK_N * 128;
    constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = BLOCK_N, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((128u >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B needs 1024-byte alignment
    constexpr uint32_t RING = (uint32_t)STAGES * STAGE_BYTES;
    constexpr uint32_t EPI = (uint32_t)(BLOCK_N / 32) * TC_A_BYTES;
    constexpr uint32_t BAR_OFF = RING > EPI ? RING : EPI;
    const uint32_t bar_full = base + BAR_OFF;                          // STAGES x 8 bytes
    const uint32_t bar_empty = bar_full + 8 * STAGES;
    const uint32_t bar_acc = bar_empty + 8 * STAGES;
    const uint32_t tmem_slot = bar_acc + 8;
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));       // generic pointer to the aligned base

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    int tile = (int)blockIdx.x - (prob > 0 ? batch.tile_end[prob - 1] : 0);
    const int tq = tile % p.tiles_w; tile /= p.tiles_w;
    const int tp = tile % p.tiles_h; tile /= p.tiles_h;
    const int tnb = tile;
    const int q0 = tq * p.tw, p0 = tp * p.th, n0 = tnb * p.tn;
    const int col0 = blockIdx.y * BLOCK_N;
    const int KB = p.ntaps * p.num_cblk;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_src) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                const int t = kb / p.num_cblk, cb = kb - t * p.num_cblk;
                const uint32_t sa = base + (uint32_t)s * STAGE_BYTES, sb = sa + TC_A_BYTES;
                mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
                tma_load_4d(sa, &map_src, bar_full + 8 * s, cb * TC_BK, q0 * p.stride + p.ox[t], p0 * p.stride + p.oy[t], n0);
                tma_load_2d(sb, &map_w, bar_full + 8 * s, p.wk[t] + cb * TC_BK, col0);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (elect_one()) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = base + (uint32_t)s * STAGE_BYTES, sb = sa + TC_A_BYTES;
                const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sb);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    // advance 8 tf32 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
                    umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(bar_empty + 8 * s);
            }
            umma_commit(bar_acc);
        }
    } else {
        // ===================================================== epilogue (warps 2..5 -> TMEM lane groups 2,3,0,1)
        const int lg = warp & 3;
        const int row = lg * 32 + lane;                 // tile row == TMEM lane
        const int iw = row % p.tw, ih = (row / p.tw) % p.th, in_ = row / (p.tw * p.th);
        const int n = n0 + in_, pp = p0 + ih, qq = q0 + iw;
        const bool valid = n < p.ON && pp < p.OH && qq < p.OW;
        const int64_t pixel = ((int64_t)n * p.FH + pp * p.o_mul + p.o_offy) * p.FW + qq * p.o_mul + p.o_offx;
        mbar_wait(bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float nz = 0.f;
        if (p.epi.noise && valid) nz = __ldg(p.epi.noise_weight) * __ldg(p.epi.noise + pixel);
#pragma unroll 1
        for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ch * 32), v);
            const int colb = col0 + ch * 32;
            tc_epilogue_math(v, p.epi, colb, pixel, p.Ncol, valid, nz);
            // staging tile for this 32-column chunk: [128 rows][128 bytes], 16-byte chunks XOR-swizzled by (row & 7)
            uint8_t* stg = smem_gen + (size_t)ch * TC_A_BYTES + (size_t)row * 128;
            tc_stage_row(stg, row, v);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2 && lane == 0) {
                tma_store_4d(&map_out, base + (uint32_t)ch * TC_A_BYTES, colb, q0, p0, n0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ shared-window persistent pair kernel
// CTA pairs (tcgen05 cta_group::2, cluster 2 x 1): two consecutive pixel tiles of the same output-channel block act as one
// 256-row MMA; each CTA loads its own A tile and HALF of the B (filter) tile, the leader issues M = 256 instructions that read
// both CTAs' shared memory, accumulators live in the TMEM of both SMs.
// Shared window: the pixel tile is 16 rows x 8 columns of one image, and ONE TMA box of (16 + oy_range) x (8 + ox_range)
// pixels (18 x 10 for a 3x3 filter, 22.5 KB per 32-channel block) serves ALL taps: tap (r, s) is the same shared-memory
// window read through a UMMA descriptor whose start address is shifted by (r * ww + s) 128-byte pixel rows and whose 8-row
// atoms are `ww` rows apart (SBO = ww * 128 B).  The 128-byte swizzle is a function of the absolute shared-memory address, so
// a row-shifted start needs no re-layout.  L2 -> SM traffic per channel block drops from 9 x 16 KB (A) to 22.5 KB.
// Persistent: one wave of CTA pairs loops over (pixel-tile pair, channel block) work items; TMEM, barriers and tensor-map
// prefetches are set up once and the mbarrier rings keep running across tiles.
constexpr int TC3_NA = 2;                           // A-window ring slots
constexpr int TC3_ASLOT = 23 * 1024;                // >= 18 * 10 * 128 B, multiple of 1024

// Overlapped epilogue: the epilogue owns two 16 KB staging buffers and the accumulators are double-buffered in TMEM
// (2 x BLOCK_N columns, two CTAs per SM), so a CTA loads and multiplies tile i + 1 while tile i streams out — what the
// memory-bound layers need (1x1 convs, 32/64-channel layers, the 1- and 2-tap parity classes of a stride-2 data gradient),
// where the stores ARE the critical path, and what lets the 3x3 layers keep the tensor pipe fed across tile boundaries.
// Timing experiments on 128 -> 128 at 256^2 x 32 (profiles/r2_tc5_time_split.txt): whole kernel 0.794 ms; without the output
// stores 0.691; loads + epilogue without any MMA 0.492; the MMAs alone need >= 0.64 ms at the 1.6 GHz the SMs hold under this
// load — the kernel sits at ~80 % of the tensor-pipe bound, the rest is imperfect overlap of the three engines.
template <int BLOCK_N> constexpr int tc5_nb() { return BLOCK_N >= 128 ? 4 : (BLOCK_N >= 64 ? 8 : 12); }   // B ring: 32 KB at N = 128 / 64, 24 KB at 32
                                                                     // (12 x 2 KB: a whole 3x3 x 32-channel filter slice, see b_resident)
constexpr int TC5_NSTG = 2;                          // dedicated 16 KB output staging buffers
// Template parameters beyond the column-block width: NACC accumulator buffers in TMEM, NA window slots, NB filter-tile slots,
// CTAS resident CTAs per SM.  Instantiated as <N, 2, 2, tc5_nb<N>(), 2>.  (Measured and dropped, round 2: one CTA per SM
// pooling the SM's shared memory with 4 accumulators and 4 + 8 slots, <128, 4, 4, 8, 1>: 757 vs 779 TFLOP/s on
// 128 -> 128 at 256^2; 256-column blocks at one CTA per SM, <256, 2, 4, 5, 1>: +3 % on 256 -> 256 / 512 -> 512, -5 % on the
// 32^2 maps, no change of the training step.)
template <int BLOCK_N, int NA, int NB>
constexpr size_t tc5_smem_bytes() {
    return (size_t)NB * (BLOCK_N / 2) * 128 + (size_t)NA * TC3_ASLOT + (size_t)TC5_NSTG * TC_A_BYTES + 1024 /*alignment*/ +
           8 * (size_t)(2 * NA + 2 * NB + 4 * 2) + 64 /*barriers incl. up to 4 accumulator pairs, TMEM slot*/;
}

template <int BLOCK_N, int NACC, int NA, int NB, int CTAS>
__global__ void __launch_bounds__(TC_THREADS, CTAS)
conv_tc5_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_out, const TcParams p, const int n_blocks, const int total_work) {
    constexpr int B_HALF_BYTES = (BLOCK_N / 2) * 128;
    constexpr int TC3_NB = NB;
    constexpr int TC3_NA = NA;                                          // (shadows the file-level constant inside this kernel)
    constexpr uint32_t B_RING = (uint32_t)TC3_NB * B_HALF_BYTES;
    constexpr uint32_t RING0 = B_RING + (uint32_t)TC3_NA * TC3_ASLOT;
    constexpr uint32_t STG_OFF = RING0;                                  // staging follows the rings (1024-byte aligned)
    constexpr uint32_t RING = RING0 + (uint32_t)TC5_NSTG * TC_A_BYTES;
    constexpr uint32_t ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr uint32_t TMEM_COLS = NACC * ACC_COLS;                      // NACC accumulator buffers
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((256u >> 4) << 24);
    constexpr int NCHUNK = BLOCK_N / 32;
    static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0 && (B_RING % 1024u) == 0 && (TC3_ASLOT % 1024) == 0, "tc5: layout");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_ring = base + B_RING;
    const uint32_t bar_fullA = base + RING;
    const uint32_t bar_emptyA = bar_fullA + 8 * TC3_NA;
    const uint32_t bar_fullB = bar_emptyA + 8 * TC3_NA;
    const uint32_t bar_emptyB = bar_fullB + 8 * TC3_NB;
    const uint32_t bar_acc = bar_emptyB + 8 * TC3_NB;              // [NACC]: accumulator buffer b is complete
    const uint32_t bar_tmem_empty = bar_acc + 8 * NACC;            // [NACC]: leader's copy in use, both CTAs' epilogues arrive on it
    const uint32_t tmem_slot = bar_tmem_empty + 8 * NACC;
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const uint32_t a_bytes = (uint32_t)(p.ww * p.wh) * 128u;
    // work item -> (pixel tile of this CTA, channel block); the channel block is the fast index
    auto decode = [&](int work, int& q0, int& p0, int& n0, int& col0) {
        const int nblk = work % n_blocks;
        int tile = (work / n_blocks) * 2 + (int)rank;
        const int tq = tile % p.tiles_w; tile /= p.tiles_w;
        const int tp = tile % p.tiles_h; tile /= p.tiles_h;
        q0 = tq * p.tw; p0 = tp * p.th; n0 = tile; col0 = nblk * BLOCK_N;
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_src) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
        for (int s = 0; s < TC3_NA; ++s) { mbar_init(bar_fullA + 8 * s, 1); mbar_init(bar_emptyA + 8 * s, 1); }
        for (int s = 0; s < TC3_NB; ++s) { mbar_init(bar_fullB + 8 * s, 1); mbar_init(bar_emptyB + 8 * s, 1); }
        for (int s = 0; s < NACC; ++s) { mbar_init(bar_acc + 8 * s, 1); mbar_init(bar_tmem_empty + 8 * s, 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (elect_one()) {
            int ia = 0, ib = 0, it = 0;
            if (p.b_resident && cluster_id < total_work) {
                // narrow layers / 1x1 convs: the filter slice of this CTA's column block is small enough to live in the B ring for
                // the whole kernel (the launcher makes every work item of a cluster use the same column block): ONE barrier,
                // ntaps x num_cblk tile loads, then per pixel tile only the activation window travels
                int q0, p0, n0, col0;
                decode(cluster_id, q0, p0, n0, col0);
                if (leader) mbar_expect_tx(bar_fullB, 2u * (uint32_t)(p.ntaps * p.num_cblk) * B_HALF_BYTES);
                for (int cb = 0; cb < p.num_cblk; ++cb)
                    for (int t = 0; t < p.ntaps; ++t)
                        tma2_load_2d(base + (uint32_t)(cb * p.ntaps + t) * B_HALF_BYTES, &map_w, bar_fullB, p.wk[t] + cb * TC_BK,
                                     col0 + (int)rank * (BLOCK_N / 2));
            }
            for (int work = cluster_id; work < total_work; work += num_clusters, ++it) {
                int q0, p0, n0, col0;
                decode(work, q0, p0, n0, col0);
                // (the output staging has its own shared memory: the loads of the next tile never wait for the epilogue)
                for (int cb = 0; cb < p.num_cblk; ++cb, ++ia) {
                    const int sa = ia % TC3_NA;
                    mbar_wait(bar_emptyA + 8 * sa, (((uint32_t)(ia / TC3_NA)) & 1u) ^ 1u);
                    if (leader) mbar_expect_tx(bar_fullA + 8 * sa, 2 * a_bytes);
                    tma2_load_4d(a_ring + (uint32_t)sa * TC3_ASLOT, &map_src, bar_fullA + 8 * sa, cb * TC_BK, q0 + p.ox_min, p0 + p.oy_min, n0);
                    if (p.b_resident) continue;
                    for (int t = 0; t < p.ntaps; ++t, ++ib) {
                        const int sb = ib % TC3_NB;
                        mbar_wait(bar_emptyB + 8 * sb, (((uint32_t)(ib / TC3_NB)) & 1u) ^ 1u);
                        if (leader) mbar_expect_tx(bar_fullB + 8 * sb, 2 * B_HALF_BYTES);
                        // (per-sample filters: both tiles of a pair lie in the same image — the launcher checks that the number
                        // of tiles per image is even — so the two CTAs agree on n0)
                        tma2_load_2d(base + (uint32_t)sb * B_HALF_BYTES, &map_w, bar_fullB + 8 * sb, p.wk[t] + cb * TC_BK,
                                     col0 + (int)rank * (BLOCK_N / 2) + n0 * p.w_nstride);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (leader CTA only)
        if (leader && elect_one()) {
            const uint64_t sbo = (uint64_t)((uint32_t)(p.ww * 128) >> 4) << 32;        // 8-row atoms are one window row apart
            int ia = 0, ib = 0, it = 0;
            if (p.b_resident && cluster_id < total_work) mbar_wait(bar_fullB, 0);        // the resident filter slice has landed
            for (int work = cluster_id; work < total_work; work += num_clusters, ++it) {
            // accumulator buffer it % NACC: both CTAs' epilogues must have drained its previous tile (it - NACC)
            const uint32_t buf = (uint32_t)(it % NACC);
            if (it >= NACC) { mbar_wait(bar_tmem_empty + 8 * buf, (uint32_t)((it / NACC) - 1) & 1u); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
            const uint32_t tmem_acc = tmem_base + buf * ACC_COLS;
            int tstep = 0;
            for (int cb = 0; cb < p.num_cblk; ++cb, ++ia) {
                const int sa = ia % TC3_NA;
                mbar_wait(bar_fullA + 8 * sa, ((uint32_t)(ia / TC3_NA)) & 1u);
                const uint32_t a0 = a_ring + (uint32_t)sa * TC3_ASLOT;
                for (int t = 0; t < p.ntaps; ++t, ++ib, ++tstep) {
                    const int sb = p.b_resident ? cb * p.ntaps + t : ib % TC3_NB;
                    if (!p.b_resident) mbar_wait(bar_fullB + 8 * sb, ((uint32_t)(ib / TC3_NB)) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // A descriptor: start = window row arow[t]; same 128B-swizzle K-major layout, SBO = ww * 128 B
                    const uint32_t aaddr = a0 + (uint32_t)p.arow[t] * 128u;
                    uint64_t da = (uint64_t)((aaddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | sbo | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
                    const uint64_t db = make_desc_sw128(base + (uint32_t)sb * B_HALF_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)
                        umma2_tf32(tmem_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (tstep > 0 || k > 0) ? 1u : 0u);
                    if (!p.b_resident) umma2_commit(bar_emptyB + 8 * sb);
                }
                umma2_commit(bar_emptyA + 8 * sa);
            }
            umma2_commit(bar_acc + 8 * buf);
            }
        }
    } else {
        // ===================================================== epilogue (identical to the pair kernel)
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const int iw = row % p.tw, ih = row / p.tw;
        int it = 0, gch = 0;         // gch: running chunk count -> staging buffer and bulk-group bookkeeping across tiles
        for (int work = cluster_id; work < total_work; work += num_clusters, ++it) {
        int q0, p0, n0, col0;
        decode(work, q0, p0, n0, col0);
        const uint32_t buf = (uint32_t)(it % NACC);
        const int n = n0, pp = p0 + ih, qq = q0 + iw;
        const bool valid = n < p.ON && pp < p.OH && qq < p.OW;
        const int64_t pixel = ((int64_t)n * p.FH + pp * p.o_mul + p.o_offy) * p.FW + qq * p.o_mul + p.o_offx;
        mbar_wait(bar_acc + 8 * buf, (uint32_t)(it / NACC) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float nz = 0.f;
        if (p.epi.noise && valid) nz = __ldg(p.epi.noise_weight) * __ldg(p.epi.noise + pixel);
#pragma unroll 1
        for (int ch = 0; ch < NCHUNK; ++ch, ++gch) {
            if (gch >= TC5_NSTG) {
                // the bulk store that last read this staging buffer (TC5_NSTG chunks ago) must have finished reading it
                if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(TC5_NSTG - 1) : "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            float v[32];
            tmem_ld32(tmem_base + buf * ACC_COLS + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ch * 32), v);
            const int colb = col0 + ch * 32;
            tc_epilogue_math(v, p.epi, colb, pixel, p.Ncol, valid, nz);
            const uint32_t stg_off = STG_OFF + (uint32_t)(gch % TC5_NSTG) * TC_A_BYTES;
            uint8_t* stg = smem_gen + stg_off + (size_t)row * 128;
            tc_stage_row(stg, row, v);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (ch == NCHUNK - 1) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2 && lane == 0) {
                tma_store_4d(&map_out, base + stg_off, colb, q0, p0, n0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (ch == NCHUNK - 1) {
                    // all 128 epilogue threads have finished reading TMEM: tell the leader's MMA warp (remote for rank 1)
                    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"((bar_tmem_empty + 8 * buf) & kPeerBitMask) : "memory");
                }
            }
        }
        }
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // before the CTA retires
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ conv_tc5 over the parity classes of a stride-2 data gradient
// The stride-2 data gradient (and the generator's transposed convolution) is four dense stride-1 problems — the parity classes
// of the output — over the SAME input.  conv_tc5m_kernel is conv_tc5_kernel whose work items carry a GROUP of ACCS classes:
// one input window per channel block (the union window of all taps) feeds the taps of every class of the group, each class
// accumulating into its own 128-column TMEM accumulator; each class has its own tap list, output sub-grid offset and output
// tensor map.
//   ACCS = 1, two CTAs per SM: four work items per pixel tile, one per class (1-, 2-, 2- and 4-tap K loops for a 3x3 filter).
//   ACCS = 2, one CTA per SM owning all 512 TMEM columns (2 buffers x 2 accumulators): two work items per pixel tile, the
//     4-tap class paired with the 1-tap class and the two 2-tap classes together — 5 and 4 taps per window load instead of
//     4 / 2 / 2 / 1.  L2 -> SM bytes per MMA clock: (4 x 19.6 KB windows + 9 x 8 KB filter tiles) per 9 taps and two CTAs per SM
//     = 130 B/clk/SM at ACCS = 1, (2 x 19.6 + 72) KB per 9 taps = 48 B/clk/SM at ACCS = 2.
// (A variant keeping all four class accumulators in TMEM per window, 4 x 64 columns at one CTA per SM, was measured in
// round 2 and dropped: 64-column MMAs are shared-memory bound; 333 vs 416 TFLOP/s, profiles/r2_prof_s2_dgrad_tc7.txt.)
constexpr int TC_MAX_CLS = 4;
constexpr int TC_GRP_TAPS = 8;
struct TcOutMaps { CUtensorMap m[TC_MAX_CLS]; };
struct TcGroups {
    int ngrp;                                   // work items per pixel-tile pair and column block
    int ntaps[TC_MAX_CLS];                      // taps of group g (all classes of the group)
    int wk[TC_MAX_CLS][TC_GRP_TAPS];            // K offset of the tap's filter slice
    unsigned short arow[TC_MAX_CLS][TC_GRP_TAPS];   // first window row of the tap
    unsigned char acc[TC_MAX_CLS][TC_GRP_TAPS];     // accumulator of the group the tap adds into
    unsigned char first[TC_MAX_CLS][TC_GRP_TAPS];   // 1: the tap opens its accumulator (overwrites instead of accumulating)
    int cls[TC_MAX_CLS][2];                     // class (output map, sub-grid offset) behind accumulator a of group g
    int o_offy[TC_MAX_CLS], o_offx[TC_MAX_CLS];
};

// ring depths: ACCS = 1 as conv_tc5_kernel (two CTAs per SM); ACCS = 2 pools the SM's shared memory: 4 window slots of 20 KB
// (17 x 9 pixels — a 3x3 filter's classes; the launcher checks), 8 filter tiles, 4 staging buffers = 208 KB
template <int ACCS> constexpr int tc5m_na() { return ACCS == 1 ? 2 : 4; }
template <int ACCS> constexpr int tc5m_aslot() { return ACCS == 1 ? TC3_ASLOT : 20 * 1024; }
template <int ACCS> constexpr int tc5m_nb() { return ACCS == 1 ? 4 : 8; }
template <int ACCS> constexpr int tc5m_nstg() { return ACCS == 1 ? 2 : 4; }
template <int ACCS>
constexpr size_t tc5m_smem_bytes() {
    return (size_t)tc5m_nb<ACCS>() * 64 * 128 + (size_t)tc5m_na<ACCS>() * tc5m_aslot<ACCS>() + (size_t)tc5m_nstg<ACCS>() * TC_A_BYTES + 1024 +
           8 * (size_t)(2 * tc5m_na<ACCS>() + 2 * tc5m_nb<ACCS>() + 4) + 64;
}

template <int ACCS>
__global__ void __launch_bounds__(64 + 128 * ACCS, ACCS == 1 ? 2 : 1)
conv_tc5m_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ TcOutMaps outs, const TcParams p, const __grid_constant__ TcGroups grp, const int n_blocks,
                 const int total_work) {
    constexpr int BLOCK_N = 128;
    constexpr int B_HALF_BYTES = (BLOCK_N / 2) * 128;
    constexpr int TC3_NB = tc5m_nb<ACCS>();
    constexpr int TC3_NA = tc5m_na<ACCS>();
    constexpr int NSTG = tc5m_nstg<ACCS>();
    constexpr uint32_t ASLOT = (uint32_t)tc5m_aslot<ACCS>();
    constexpr uint32_t B_RING = (uint32_t)TC3_NB * B_HALF_BYTES;
    constexpr uint32_t RING0 = B_RING + (uint32_t)TC3_NA * ASLOT;
    constexpr uint32_t STG_OFF = RING0;                                  // staging follows the rings (1024-byte aligned)
    constexpr uint32_t RING = RING0 + (uint32_t)NSTG * TC_A_BYTES;
    constexpr uint32_t ACC_COLS = (uint32_t)(ACCS * BLOCK_N);            // one accumulator buffer = ACCS class accumulators
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;                         // two accumulator buffers
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((256u >> 4) << 24);
    constexpr int NCHUNK = BLOCK_N / 32;
    static_assert(TMEM_COLS <= 512 && (B_RING % 1024u) == 0 && (ASLOT % 1024u) == 0, "tc5m: layout");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_ring = base + B_RING;
    const uint32_t bar_fullA = base + RING;
    const uint32_t bar_emptyA = bar_fullA + 8 * TC3_NA;
    const uint32_t bar_fullB = bar_emptyA + 8 * TC3_NA;
    const uint32_t bar_emptyB = bar_fullB + 8 * TC3_NB;
    const uint32_t bar_acc = bar_emptyB + 8 * TC3_NB;              // [2]: accumulator buffer b is complete
    const uint32_t bar_tmem_empty = bar_acc + 16;                  // [2]: leader's copy in use, both CTAs' epilogues arrive on it
    const uint32_t tmem_slot = bar_tmem_empty + 16;
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    // contiguous share of the work items per cluster: the groups of one pixel tile differ in their tap counts, a strided
    // assignment would hand some clusters only the long ones; and consecutive items re-read the same window from L2
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int work_begin = (int)(((int64_t)cluster_id * total_work) / num_clusters);
    const int work_end = (int)(((int64_t)(cluster_id + 1) * total_work) / num_clusters);
    const uint32_t a_bytes = (uint32_t)(p.ww * p.wh) * 128u;
    // work item -> (pixel tile of this CTA, class group, channel block); channel block fastest, then the group
    auto decode = [&](int work, int& q0, int& p0, int& n0, int& col0, int& g) {
        const int nblk = work % n_blocks;
        const int rest = work / n_blocks;
        g = rest % grp.ngrp;
        int tile = (rest / grp.ngrp) * 2 + (int)rank;
        const int tq = tile % p.tiles_w; tile /= p.tiles_w;
        const int tp = tile % p.tiles_h; tile /= p.tiles_h;
        q0 = tq * p.tw; p0 = tp * p.th; n0 = tile; col0 = nblk * BLOCK_N;
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_src) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int c = 0; c < TC_MAX_CLS; ++c) asm volatile("prefetch.tensormap [%0];" ::"l"(&outs.m[c]) : "memory");
        for (int s = 0; s < TC3_NA; ++s) { mbar_init(bar_fullA + 8 * s, 1); mbar_init(bar_emptyA + 8 * s, 1); }
        for (int s = 0; s < TC3_NB; ++s) { mbar_init(bar_fullB + 8 * s, 1); mbar_init(bar_emptyB + 8 * s, 1); }
        mbar_init(bar_acc, 1); mbar_init(bar_acc + 8, 1);
        mbar_init(bar_tmem_empty, 2 * ACCS); mbar_init(bar_tmem_empty + 8, 2 * ACCS);      // both CTAs x ACCS epilogue groups
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (elect_one()) {
            int ia = 0, ib = 0;
            for (int work = work_begin; work < work_end; ++work) {
                int q0, p0, n0, col0, g;
                decode(work, q0, p0, n0, col0, g);
                const int ntaps = grp.ntaps[g];
                for (int cb = 0; cb < p.num_cblk; ++cb, ++ia) {
                    const int sa = ia % TC3_NA;
                    mbar_wait(bar_emptyA + 8 * sa, (((uint32_t)(ia / TC3_NA)) & 1u) ^ 1u);
                    if (leader) mbar_expect_tx(bar_fullA + 8 * sa, 2 * a_bytes);
                    tma2_load_4d(a_ring + (uint32_t)sa * ASLOT, &map_src, bar_fullA + 8 * sa, cb * TC_BK, q0 + p.ox_min, p0 + p.oy_min, n0);
                    for (int t = 0; t < ntaps; ++t, ++ib) {
                        const int sb = ib % TC3_NB;
                        mbar_wait(bar_emptyB + 8 * sb, (((uint32_t)(ib / TC3_NB)) & 1u) ^ 1u);
                        if (leader) mbar_expect_tx(bar_fullB + 8 * sb, 2 * B_HALF_BYTES);
                        tma2_load_2d(base + (uint32_t)sb * B_HALF_BYTES, &map_w, bar_fullB + 8 * sb, grp.wk[g][t] + cb * TC_BK,
                                     col0 + (int)rank * (BLOCK_N / 2));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (leader CTA only)
        if (leader && elect_one()) {
            const uint64_t sbo = (uint64_t)((uint32_t)(p.ww * 128) >> 4) << 32;        // 8-row atoms are one window row apart
            int ia = 0, ib = 0, it = 0;
            for (int work = work_begin; work < work_end; ++work, ++it) {
            // accumulator buffer (it & 1): both CTAs' epilogues must have drained its previous item (it - 2)
            const uint32_t buf = (uint32_t)it & 1u;
            if (it > 1) { mbar_wait(bar_tmem_empty + 8 * buf, (uint32_t)((it >> 1) - 1) & 1u); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
            const int g = (work / n_blocks) % grp.ngrp;
            const int ntaps = grp.ntaps[g];
            for (int cb = 0; cb < p.num_cblk; ++cb, ++ia) {
                const int sa = ia % TC3_NA;
                mbar_wait(bar_fullA + 8 * sa, ((uint32_t)(ia / TC3_NA)) & 1u);
                const uint32_t a0 = a_ring + (uint32_t)sa * ASLOT;
                for (int t = 0; t < ntaps; ++t, ++ib) {
                    const int sb = ib % TC3_NB;
                    mbar_wait(bar_fullB + 8 * sb, ((uint32_t)(ib / TC3_NB)) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // A descriptor: start = the tap's window row; same 128B-swizzle K-major layout, SBO = ww * 128 B
                    const uint32_t aaddr = a0 + (uint32_t)grp.arow[g][t] * 128u;
                    uint64_t da = (uint64_t)((aaddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | sbo | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
                    const uint64_t db = make_desc_sw128(base + (uint32_t)sb * B_HALF_BYTES);
                    const uint32_t tmem_acc = tmem_base + buf * ACC_COLS + (uint32_t)grp.acc[g][t] * BLOCK_N;
                    const bool opens = cb == 0 && grp.first[g][t] != 0;
                    if (!(p.debug & 4)) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)
                        umma2_tf32(tmem_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (!opens || k > 0) ? 1u : 0u);
                    }
                    umma2_commit(bar_emptyB + 8 * sb);
                }
                umma2_commit(bar_emptyA + 8 * sa);
            }
            umma2_commit(bar_acc + 8 * buf);
            }
        }
    } else {
        // ===================================================== epilogue: one group of four warps PER class accumulator of the item
        // (as in conv_tc5_kernel; ncu's stall sampling of the single-group version showed the epilogue warps busy ~90 % of the
        // kernel — two 64 KB tiles per item and CTA against 4-5 taps of MMAs — and the tensor pipe waiting for free accumulators)
        const int eg = (warp - 2) >> 2;                  // epilogue group = accumulator within the item
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const int iw = row % p.tw, ih = row / p.tw;
        const bool boss = ((warp - 2) & 3) == 0 && lane == 0;          // the group's store-issuing thread
        const uint32_t barid = 1u + (uint32_t)eg;
        constexpr int GS = NSTG / ACCS;                  // staging buffers per group
        int it = 0, gch = 0;         // gch: the group's running chunk count -> staging buffer and bulk-group bookkeeping across tiles
        for (int work = work_begin; work < work_end; ++work, ++it) {
        int q0, p0, n0, col0, g;
        decode(work, q0, p0, n0, col0, g);
        const uint32_t buf = (uint32_t)it & 1u;
        const int n = n0, pp = p0 + ih, qq = q0 + iw;
        const bool valid = n < p.ON && pp < p.OH && qq < p.OW;
        mbar_wait(bar_acc + 8 * buf, (uint32_t)(it >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int c = grp.cls[g][eg];
        const int64_t pixel = ((int64_t)n * p.FH + pp * p.o_mul + grp.o_offy[c]) * p.FW + qq * p.o_mul + grp.o_offx[c];
        float nz = 0.f;
        if (p.epi.noise && valid) nz = __ldg(p.epi.noise_weight) * __ldg(p.epi.noise + pixel);
#pragma unroll 1
        for (int ch = 0; ch < NCHUNK; ++ch, ++gch) {
            const bool last = ch == NCHUNK - 1;
            if (gch >= GS) {
                // the bulk store that last read this staging buffer (GS chunks ago) must have finished reading it
                if (boss) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(GS - 1) : "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
            }
            float v[32];
            const int colb = col0 + ch * 32;
            const uint32_t stg_off = STG_OFF + (uint32_t)(eg * GS + gch % GS) * TC_A_BYTES;
            if (!(p.debug & 2)) {
            tmem_ld32(tmem_base + buf * ACC_COLS + ((uint32_t)(lg * 32) << 16) + (uint32_t)(eg * BLOCK_N + ch * 32), v);
            tc_epilogue_math(v, p.epi, colb, pixel, p.Ncol, valid, nz);
            uint8_t* stg = smem_gen + stg_off + (size_t)row * 128;
            tc_stage_row(stg, row, v);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (last) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
            if (boss) {
                if (!(p.debug & 3)) tma_store_4d(&outs.m[c], base + stg_off, colb, q0, p0, n0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (last) {
                    // this group's 128 threads have finished reading TMEM: tell the leader's MMA warp (remote for rank 1)
                    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"((bar_tmem_empty + 8 * buf) & kPeerBitMask) : "memory");
                }
            }
        }
        }
        if (boss) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // before the CTA retires
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ persistent per-tap pair kernel
// Stride-2 fprop and maps too small for the shared window: CTA pairs, one 4-D TMA box per tap (the conv stride is the TMA
// element stride), a persistent loop over (tile pair, column block) work items, two accumulator buffers in TMEM and dedicated
// output staging, so loads and MMAs of tile i + 1 overlap the epilogue of tile i.
// BLOCK_N = 128: 3 x (16 KB A + 8 KB half-B) + 2 x 16 KB staging = 104 KB, two CTAs per SM.
// BLOCK_N = 256: ONE CTA per SM owning all 512 TMEM columns (2 x 256), a 5-deep ring and two epilogue groups: 5 x (16 KB + 16 KB) + 4 x 16 KB = 224 KB.
// The per-tap kernels are bound by the L2 -> SM feed (ncu, profiles/r2_prof_s2_fprop_tc6.txt: lts 58 %, tensor pipe 53 %):
// 256 columns halve the A bytes per MAC and the deeper ring covers the L2 latency that 3 stages x 256 clk cannot.
template <int BLOCK_N>
constexpr int tc6_stages() { return BLOCK_N >= 256 ? 5 : 3; }

// epilogue groups of four warps: one per 128 output columns (BLOCK_N = 256: two groups, each with its own two staging buffers
// and named barrier — a single group drained 8 x 16 KB per item against a K loop that is only twice as long as at 128 columns)
template <int BLOCK_N> constexpr int tc6_groups() { return BLOCK_N / 128; }
template <int BLOCK_N> constexpr int tc6_threads() { return 64 + 128 * tc6_groups<BLOCK_N>(); }

template <int BLOCK_N>
constexpr size_t tc6_smem_bytes() {
    return (size_t)tc6_stages<BLOCK_N>() * (TC_A_BYTES + (BLOCK_N / 2) * 128) + (size_t)(tc6_groups<BLOCK_N>() * TC5_NSTG) * TC_A_BYTES + 1024 + 256;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(tc6_threads<BLOCK_N>(), BLOCK_N >= 256 ? 1 : 2)
conv_tc6_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_out, const TcParams p, const int n_blocks, const int total_work) {
    constexpr int STAGES = tc6_stages<BLOCK_N>();
    constexpr int B_HALF_BYTES = (BLOCK_N / 2) * 128;
    constexpr int STAGE_BYTES = TC_A_BYTES + B_HALF_BYTES;                 // per CTA
    constexpr uint32_t ACC_COLS = BLOCK_N;
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;                           // two accumulator buffers
    static_assert(BLOCK_N == 128 || BLOCK_N == 256, "tc6: 2 x 128 TMEM columns at two CTAs per SM, or 2 x 256 at one");
    // D fp32, A/B tf32 K-major, N = BLOCK_N, M = 256 (128 rows from each CTA)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((256u >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t RING = (uint32_t)STAGES * STAGE_BYTES;
    constexpr uint32_t STG_OFF = RING;                                     // dedicated staging behind the ring
    constexpr int EG = tc6_groups<BLOCK_N>();
    constexpr uint32_t BAR_OFF = RING + (uint32_t)(EG * TC5_NSTG) * TC_A_BYTES;
    static_assert(RING % 1024u == 0, "tc6: staging must stay 1024-byte aligned");
    constexpr int NCHUNK = BLOCK_N / 32 / EG;                              // 32-column chunks per epilogue group
    const uint32_t bar_full = base + BAR_OFF;
    const uint32_t bar_empty = bar_full + 8 * STAGES;
    const uint32_t bar_acc = bar_empty + 8 * STAGES;              // [2]
    const uint32_t bar_tmem_empty = bar_acc + 16;                 // [2], leader's copy in use
    const uint32_t tmem_slot = bar_tmem_empty + 16;
    uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    // persistent: one wave of CTA pairs loops over work items = (pair of consecutive pixel tiles, output-channel block),
    // the channel block being the fast index so that the pairs re-reading one activation tile run together (L2 hits)
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    auto decode = [&](int work, int& q0, int& p0, int& n0, int& col0) {
        const int nblk = work % n_blocks;
        int tile = (work / n_blocks) * 2 + (int)rank;
        const int tq = tile % p.tiles_w; tile /= p.tiles_w;
        const int tp = tile % p.tiles_h; tile /= p.tiles_h;
        q0 = tq * p.tw; p0 = tp * p.th; n0 = tile * p.tn; col0 = nblk * BLOCK_N;      // beyond the batch for a padding tile: all OOB
    };
    const int KB = p.ntaps * p.num_cblk;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_src) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1); mbar_init(bar_acc + 8, 1);
        mbar_init(bar_tmem_empty, 2 * EG); mbar_init(bar_tmem_empty + 8, 2 * EG);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                      // both CTAs' barriers exist before any remote complete_tx / commit arrives
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (elect_one()) {
            int kg = 0;                                       // ring position runs on across tiles
            for (int work = cluster_id; work < total_work; work += num_clusters) {
            int q0, p0, n0, col0;
            decode(work, q0, p0, n0, col0);
            for (int kb = 0; kb < KB; ++kb, ++kg) {
                const int s = kg % STAGES;
                const uint32_t ph = (uint32_t)(kg / STAGES) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                const int t = kb / p.num_cblk, cb = kb - t * p.num_cblk;
                const uint32_t sa = base + (uint32_t)s * STAGE_BYTES, sb = sa + TC_A_BYTES;
                if (leader) mbar_expect_tx(bar_full + 8 * s, 2 * STAGE_BYTES);     // bytes of both CTAs land on the leader's barrier
                tma2_load_4d(sa, &map_src, bar_full + 8 * s, cb * TC_BK, q0 * p.stride + p.ox[t], p0 * p.stride + p.oy[t], n0);
                tma2_load_2d(sb, &map_w, bar_full + 8 * s, p.wk[t] + cb * TC_BK, col0 + (int)rank * (BLOCK_N / 2));
            }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (leader CTA only)
        if (leader && elect_one()) {
            int kg = 0, it = 0;
            for (int work = cluster_id; work < total_work; work += num_clusters, ++it) {
            // accumulator buffer (it & 1): both CTAs' epilogues must have drained its previous tile (it - 2)
            const uint32_t buf = (uint32_t)it & 1u;
            if (it > 1) { mbar_wait(bar_tmem_empty + 8 * buf, (uint32_t)((it >> 1) - 1) & 1u); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
            const uint32_t tmem_acc = tmem_base + buf * ACC_COLS;
            for (int kb = 0; kb < KB; ++kb, ++kg) {
                const int s = kg % STAGES;
                const uint32_t ph = (uint32_t)(kg / STAGES) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = base + (uint32_t)s * STAGE_BYTES, sb = sa + TC_A_BYTES;
                const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sb);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k)
                    umma2_tf32(tmem_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                umma2_commit(bar_empty + 8 * s);          // frees stage s in both CTAs
            }
            umma2_commit(bar_acc + 8 * buf);              // this buffer's accumulators complete in both CTAs
            }
        }
    } else {
        // ===================================================== epilogue: group eg of four warps owns columns [128 eg, 128 eg + 128)
        const int eg = (warp - 2) >> 2;
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const int iw = row % p.tw, ih = (row / p.tw) % p.th, in_ = row / (p.tw * p.th);
        const bool boss = ((warp - 2) & 3) == 0 && lane == 0;
        const uint32_t barid = 1u + (uint32_t)eg;
        int it = 0, gch = 0;
        for (int work = cluster_id; work < total_work; work += num_clusters, ++it) {
        int q0, p0, n0, col0;
        decode(work, q0, p0, n0, col0);
        const uint32_t buf = (uint32_t)it & 1u;
        const int n = n0 + in_, pp = p0 + ih, qq = q0 + iw;
        const bool valid = n < p.ON && pp < p.OH && qq < p.OW;
        const int64_t pixel = ((int64_t)n * p.FH + pp * p.o_mul + p.o_offy) * p.FW + qq * p.o_mul + p.o_offx;
        mbar_wait(bar_acc + 8 * buf, (uint32_t)(it >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float nz = 0.f;
        if (p.epi.noise && valid) nz = __ldg(p.epi.noise_weight) * __ldg(p.epi.noise + pixel);
#pragma unroll 1
        for (int ch = 0; ch < NCHUNK; ++ch, ++gch) {
            if (gch >= TC5_NSTG) {
                // the bulk store that last read this staging buffer (TC5_NSTG chunks ago) must have finished reading it
                if (boss) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(TC5_NSTG - 1) : "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
            }
            float v[32];
            tmem_ld32(tmem_base + buf * ACC_COLS + ((uint32_t)(lg * 32) << 16) + (uint32_t)(eg * 128 + ch * 32), v);
            const int colb = col0 + eg * 128 + ch * 32;
            tc_epilogue_math(v, p.epi, colb, pixel, p.Ncol, valid, nz);
            const uint32_t stg_off = STG_OFF + (uint32_t)(eg * TC5_NSTG + gch % TC5_NSTG) * TC_A_BYTES;
            uint8_t* stg = smem_gen + stg_off + (size_t)row * 128;
            tc_stage_row(stg, row, v);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (ch == NCHUNK - 1) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
            if (boss) {
                tma_store_4d(&map_out, base + stg_off, colb, q0, p0, n0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (ch == NCHUNK - 1)      // this group's 128 threads are done with the accumulator buffer
                    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"((bar_tmem_empty + 8 * buf) & kPeerBitMask) : "memory");
            }
        }
        }
        if (boss) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                      // nobody leaves (or frees TMEM) while the peer may still touch this CTA
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapSwizzle swizzle) {
    if (g_encode == nullptr && !tc_available()) return fail(SAE_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available on this device / driver");
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SAE_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SAE_OK;
}

struct TcProblem {
    const float* src; int SN, SH, SW, SC;      // source activation [SN,SH,SW,SC]
    const float* wmat; int Ncol, Ktot;          // filter matrix [Ncol, Ktot = ntaps*SC]
    int w_per_sample;                           // 1: wmat holds one filter matrix per image, [SN, Ncol, Ktot]
    float* out; int OH, OW;                     // output sub-grid [SN, OH, OW, Ncol] ...
    int o_mul, o_offy, o_offx, FH, FW;          // ... placed at (o_mul*p + o_offy, o_mul*q + o_offx) of the full [SN,FH,FW,Ncol]
    int stride, ntaps;
    int oy[TC_MAX_TAPS], ox[TC_MAX_TAPS], wk[TC_MAX_TAPS];
};

static bool tc_shape_ok(int src_c, int ncol, int ntaps, int stride, int ow) {
    if (src_c % 32 != 0 || ncol % 32 != 0) return false;
    if (ntaps < 1 || ntaps > TC_MAX_TAPS) return false;
    if (stride != 1 && stride != 2) return false;
    int tw = pow2_ceil(ow) < 16 ? pow2_ceil(ow) : 16;
    if (tw * stride > 256) return false;
    return true;
}

static void tc_fill_params(const TcProblem& pr, const EpiParams& e, TcParams& p) {
    p.num_cblk = pr.SC / 32;
    p.ntaps = pr.ntaps;
    p.stride = pr.stride;
    p.tw = pow2_ceil(pr.OW) < 16 ? pow2_ceil(pr.OW) : 16;
    int th = 128 / p.tw;
    if (pow2_ceil(pr.OH) < th) th = pow2_ceil(pr.OH);
    p.th = th;
    p.tn = 128 / (p.tw * p.th);
    p.tiles_w = (pr.OW + p.tw - 1) / p.tw;
    p.tiles_h = (pr.OH + p.th - 1) / p.th;
    p.tiles_n = (pr.SN + p.tn - 1) / p.tn;
    p.ON = pr.SN; p.OH = pr.OH; p.OW = pr.OW; p.Ncol = pr.Ncol; p.src_c = pr.SC;
    for (int t = 0; t < pr.ntaps; ++t) { p.oy[t] = (short)pr.oy[t]; p.ox[t] = (short)pr.ox[t]; p.wk[t] = pr.wk[t]; }
    p.o_mul = pr.o_mul; p.o_offy = pr.o_offy; p.o_offx = pr.o_offx; p.FH = pr.FH; p.FW = pr.FW;
    p.w_nstride = pr.w_per_sample ? pr.Ncol : 0;
    p.b_resident = 0;
    static int dbg = -1;
    if (dbg < 0) { const char* v = getenv("SAE_TC_DEBUG"); dbg = v ? atoi(v) : 0; }
    p.debug = dbg;
    p.epi = e;
}

static int tc_encode_maps(const TcProblem& pr, const TcParams& p, int b_rows, CUtensorMap* msrc, CUtensorMap* mw, CUtensorMap* mout) {
    {
        cuuint64_t dims[4] = {(cuuint64_t)pr.SC, (cuuint64_t)pr.SW, (cuuint64_t)pr.SH, (cuuint64_t)pr.SN};
        cuuint64_t strides[3] = {(cuuint64_t)pr.SC * 4, (cuuint64_t)pr.SW * pr.SC * 4, (cuuint64_t)pr.SH * pr.SW * pr.SC * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)(p.tw * pr.stride), (cuuint32_t)(p.th * pr.stride), (cuuint32_t)p.tn};
        cuuint32_t es[4] = {1, (cuuint32_t)pr.stride, (cuuint32_t)pr.stride, 1};
        int rc = encode_map(msrc, pr.src, 4, dims, strides, box, es);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)pr.Ktot, (cuuint64_t)pr.Ncol * (cuuint64_t)(pr.w_per_sample ? pr.SN : 1)};
        cuuint64_t strides[1] = {(cuuint64_t)pr.Ktot * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)b_rows};
        cuuint32_t es[2] = {1, 1};
        int rc = encode_map(mw, pr.wmat, 2, dims, strides, box, es);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)pr.Ncol, (cuuint64_t)pr.OW, (cuuint64_t)pr.OH, (cuuint64_t)pr.SN};
        cuuint64_t strides[3] = {(cuuint64_t)pr.o_mul * pr.Ncol * 4, (cuuint64_t)pr.o_mul * pr.FW * pr.Ncol * 4,
                                 (cuuint64_t)pr.FH * pr.FW * pr.Ncol * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn};
        cuuint32_t es[4] = {1, 1, 1, 1};
        int rc = encode_map(mout, pr.out + ((int64_t)pr.o_offy * pr.FW + pr.o_offx) * pr.Ncol, 4, dims, strides, box, es);
        if (rc) return rc;
    }
    return SAE_OK;
}

// one-tile-per-CTA launch of up to TC_MAX_BATCH problems that share the filter matrix, the column count and the epilogue
template <int BLOCK_N>
static int tc_launch_batch(const TcProblem* prs, int count, const EpiParams& e, cudaStream_t st) {
    TcBatch b;
    if (count < 1 || count > TC_MAX_BATCH) return fail(SAE_E_INVALID, "conv_tc: batch of %d problems", count);
    b.count = count;
    int tiles = 0;
    for (int i = 0; i < count; ++i) {
        tc_fill_params(prs[i], e, b.p[i]);
        CUtensorMap mw;
        int rc = tc_encode_maps(prs[i], b.p[i], BLOCK_N, &b.src[i], &mw, &b.out[i]);
        if (rc) return rc;
        if (i == 0) b.w = mw;
        else if (prs[i].wmat != prs[0].wmat || prs[i].Ncol != prs[0].Ncol || prs[i].Ktot != prs[0].Ktot)
            return fail(SAE_E_INVALID, "conv_tc: batched problems must share the filter matrix");
        tiles += b.p[i].tiles_w * b.p[i].tiles_h * b.p[i].tiles_n;
        b.tile_end[i] = tiles;
    }
    constexpr size_t smem = tc_smem_bytes<BLOCK_N>();
    static bool attr_done = false;
    if (!attr_done) {
        SAE_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    dim3 grid((unsigned)tiles, (unsigned)(prs[0].Ncol / BLOCK_N));
    conv_tc_kernel<BLOCK_N><<<grid, TC_THREADS, smem, st>>>(b);
    return check_launch("conv_tc");
}

template <int BLOCK_N>
static int tc_launch(const TcProblem& pr, const EpiParams& e, cudaStream_t st) {
    return tc_launch_batch<BLOCK_N>(&pr, 1, e, st);
}

// persistent per-tap pair launch (conv_tc6_kernel): stride-2 fprop, small maps — everything the shared window does not take
template <int BLOCK_N>
static int tc6_launch(const TcProblem& pr, const EpiParams& e, cudaStream_t st) {
    TcParams p;
    tc_fill_params(pr, e, p);
    CUtensorMap msrc, mw, mout;
    int rc = tc_encode_maps(pr, p, BLOCK_N / 2, &msrc, &mw, &mout);
    if (rc) return rc;
    constexpr size_t smem = tc6_smem_bytes<BLOCK_N>();
    static bool attr_done = false;
    if (!attr_done) {
        SAE_CUDA_TRY(cudaFuncSetAttribute(conv_tc6_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const int tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int pairs = (tiles + 1) / 2;
    const int n_blocks = pr.Ncol / BLOCK_N;
    cudaLaunchConfig_t cfg = {};
    const int total_work = pairs * n_blocks;
    // persistent: one wave of CTA pairs — two CTAs per SM at 128 columns, one at 256
    int clusters = BLOCK_N >= 256 ? sm_count() / 2 : sm_count();
    if (clusters > total_work) clusters = total_work;
    cfg.gridDim = dim3((unsigned)(clusters * 2), 1, 1);
    cfg.blockDim = dim3(tc6_threads<BLOCK_N>(), 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SAE_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_tc6_kernel<BLOCK_N>, msrc, mw, mout, p, n_blocks, total_work));
    return check_launch("conv_tc6");
}

// shared-window persistent launch (conv_tc5_kernel): tile = 16 rows x 8 columns of one image; window = tile + tap offset range;
// one wave of CTA pairs, CTAS CTAs per SM
template <int BLOCK_N, int NACC = 2, int NA = TC3_NA, int NB = tc5_nb<BLOCK_N>(), int CTAS = 2>
static int tc5_launch(const TcProblem& pr, const EpiParams& e, cudaStream_t st) {
    TcParams p;
    tc_fill_params(pr, e, p);
    p.tw = 8; p.th = 16; p.tn = 1;
    p.tiles_w = (pr.OW + p.tw - 1) / p.tw;
    p.tiles_h = (pr.OH + p.th - 1) / p.th;
    p.tiles_n = pr.SN;
    int oy_min = pr.oy[0], oy_max = pr.oy[0], ox_min = pr.ox[0], ox_max = pr.ox[0];
    for (int t = 1; t < pr.ntaps; ++t) {
        oy_min = pr.oy[t] < oy_min ? pr.oy[t] : oy_min; oy_max = pr.oy[t] > oy_max ? pr.oy[t] : oy_max;
        ox_min = pr.ox[t] < ox_min ? pr.ox[t] : ox_min; ox_max = pr.ox[t] > ox_max ? pr.ox[t] : ox_max;
    }
    p.oy_min = oy_min; p.ox_min = ox_min;
    p.ww = p.tw + (ox_max - ox_min);
    p.wh = p.th + (oy_max - oy_min);
    for (int t = 0; t < pr.ntaps; ++t) p.arow[t] = (unsigned short)((pr.oy[t] - oy_min) * p.ww + (pr.ox[t] - ox_min));
    CUtensorMap msrc, mw, mout;
    int rc = tc_encode_maps(pr, p, BLOCK_N / 2, &msrc, &mw, &mout);
    if (rc) return rc;
    {
        cuuint64_t dims[4] = {(cuuint64_t)pr.SC, (cuuint64_t)pr.SW, (cuuint64_t)pr.SH, (cuuint64_t)pr.SN};
        cuuint64_t strides[3] = {(cuuint64_t)pr.SC * 4, (cuuint64_t)pr.SW * pr.SC * 4, (cuuint64_t)pr.SH * pr.SW * pr.SC * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.ww, (cuuint32_t)p.wh, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        rc = encode_map(&msrc, pr.src, 4, dims, strides, box, es);
        if (rc) return rc;
    }
    constexpr size_t smem = tc5_smem_bytes<BLOCK_N, NA, NB>();
    static bool attr_done = false;
    if (!attr_done) {
        SAE_CUDA_TRY(cudaFuncSetAttribute(conv_tc5_kernel<BLOCK_N, NACC, NA, NB, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const int tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int n_blocks = pr.Ncol / BLOCK_N;
    const int total_work = ((tiles + 1) / 2) * n_blocks;
    int clusters = sm_count() * CTAS / 2;            // CTAS CTAs per SM, 2 CTAs per cluster
    if (clusters > total_work) clusters = total_work;
    // resident filter slice: it must fit the B ring, be shared by the batch, and every work item of a cluster must use the same
    // column block (work = cluster + i * clusters, column block = work % n_blocks)
    static int resident = -1;
    if (resident < 0) { const char* v = getenv("SAE_TC_RESIDENT_B"); resident = (v && v[0] == '0') ? 0 : 1; }
    // (measured, profiles/r2_resident_b.txt: +5 .. +13 % on the 32-channel 3x3 layers, -4 .. -9 % on 1x1 convs, whose single
    // filter tile per channel block was never the bottleneck — hence ntaps > 1)
    if (resident && !pr.w_per_sample && pr.ntaps > 1 && pr.ntaps * p.num_cblk <= NB && clusters % n_blocks == 0 && total_work >= 4 * clusters)
        p.b_resident = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * 2), 1, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SAE_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_tc5_kernel<BLOCK_N, NACC, NA, NB, CTAS>, msrc, mw, mout, p, n_blocks, total_work));
    return check_launch("conv_tc5");
}

// can the shared-window kernel take the problem?  (stride 1, a tile of 16 x 8 output pixels exists, tap offsets span <= 2)
static bool window_ok(const TcProblem& pr) {
    if (pr.stride != 1 || pr.OW < 8 || pr.OH < 16 || pr.Ncol % 32 != 0) return false;
    int oy_min = pr.oy[0], oy_max = pr.oy[0], ox_min = pr.ox[0], ox_max = pr.ox[0];
    for (int t = 1; t < pr.ntaps; ++t) {
        oy_min = pr.oy[t] < oy_min ? pr.oy[t] : oy_min; oy_max = pr.oy[t] > oy_max ? pr.oy[t] : oy_max;
        ox_min = pr.ox[t] < ox_min ? pr.ox[t] : ox_min; ox_max = pr.ox[t] > ox_max ? pr.ox[t] : ox_max;
    }
    return (oy_max - oy_min) <= 2 && (ox_max - ox_min) <= 2;
}

static int tc_dispatch(const TcProblem& pr, const EpiParams& e, cudaStream_t st) {
    // enough pixel tiles to make CTA pairs worthwhile
    if ((int64_t)pr.SN * pr.OH * pr.OW >= 2 * 128) {
        if (window_ok(pr)) {
            // every stride-1 shared-window problem: 3x3 / 1x1 fprop and dgrad at all widths, the parity classes of a stride-2
            // data gradient (measured against the one-tile kernels it replaced: profiles/r1_conv_bench_tc5.txt)
            if (pr.Ncol % 128 == 0) return tc5_launch<128>(pr, e, st);
            return pr.Ncol % 64 == 0 ? tc5_launch<64>(pr, e, st) : tc5_launch<32>(pr, e, st);
        }
        // per-tap loads (the conv stride is the TMA element stride): 256-column blocks at one CTA per SM where the layer has
        // them (565 -> 722 TFLOP/s on 128 -> 256 at 257^2 stride 2, profiles/r2_conv_bench_s2.txt), 128-column blocks otherwise
        if (pr.Ncol % 256 == 0) return tc6_launch<256>(pr, e, st);
        if (pr.Ncol % 128 == 0) return tc6_launch<128>(pr, e, st);
    }
    if (pr.Ncol % 128 == 0) return tc_launch<128>(pr, e, st);
    if (pr.Ncol % 64 == 0) return tc_launch<64>(pr, e, st);
    return tc_launch<32>(pr, e, st);
}

// Sub-rectangle [r0, r1) x [c0, c1) of a problem's output grid as a problem of its own: the output placement moves
// with it and the tap offsets absorb the shift of the origin.
static TcProblem tc_subproblem(const TcProblem& pr, int r0, int r1, int c0, int c1) {
    TcProblem s = pr;
    s.OH = r1 - r0; s.OW = c1 - c0;
    s.o_offy = pr.o_offy + pr.o_mul * r0; s.o_offx = pr.o_offx + pr.o_mul * c0;
    for (int t = 0; t < pr.ntaps; ++t) { s.oy[t] = pr.oy[t] + r0 * pr.stride; s.ox[t] = pr.ox[t] + c0 * pr.stride; }
    return s;
}

// The shared-window kernels tile the output in 16 x 8 pixel blocks.  The parity classes of a stride-2 data gradient are
// (2^k + 1)-sized in the discriminators (blurred 65 / 129 / 257 maps): one extra row and column would cost a whole extra
// row and column of mostly empty 128-pixel tiles (33 x 33 outputs -> 15 tiles instead of 8.5).  Such thin remainders are
// split off and run as their own launches, whose tiles gather the strip across images (tn > 1) instead.
static int tc_dispatch_split(const TcProblem& pr, const EpiParams& e, cudaStream_t st) {
    if (!window_ok(pr)) return tc_dispatch(pr, e, st);
    const int rem_h = pr.OH % 16, rem_w = pr.OW % 8;
    const int main_h = (rem_h >= 1 && rem_h <= 4 && pr.OH >= 32) ? pr.OH - rem_h : pr.OH;
    const int main_w = (rem_w >= 1 && rem_w <= 2 && pr.OW >= 16) ? pr.OW - rem_w : pr.OW;
    if (main_h == pr.OH && main_w == pr.OW) return tc_dispatch(pr, e, st);
    int rc = tc_dispatch(tc_subproblem(pr, 0, main_h, 0, main_w), e, st);
    if (rc) return rc;
    if (main_w < pr.OW) {                                    // right strip, full height (takes the corner)
        rc = tc_dispatch(tc_subproblem(pr, 0, pr.OH, main_w, pr.OW), e, st);
        if (rc) return rc;
    }
    if (main_h < pr.OH) rc = tc_dispatch(tc_subproblem(pr, main_h, pr.OH, 0, main_w), e, st);
    return rc;
}

static bool ptr_ok(const void* a, const void* b, const void* c) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

bool tc_fprop_eligible(const sae_conv_geom* g) {
    if (g->R * g->S > TC_MAX_TAPS) return false;
    if (g->pad_t > 100 || g->pad_l > 100) return false;
    return tc_shape_ok(g->C, g->K, g->R * g->S, g->stride, g->Q);
}

bool tc_dgrad_eligible(const sae_conv_geom* g) {
    if (g->stride != 1 && g->stride != 2) return false;
    if (g->R * g->S > TC_MAX_TAPS) return false;
    if (g->pad_t > 100 || g->pad_l > 100) return false;
    return tc_shape_ok(g->K, g->C, g->R * g->S, 1, (g->W + g->stride - 1) / g->stride);
}

int tc_fprop(const float* x, const float* w, float* y, const sae_conv_geom* g, const EpiParams& e, cudaStream_t st) {
    if (!ptr_ok(x, w, y)) return fail(SAE_E_INVALID, "conv2d_fprop(tcgen05): pointers must be 16-byte aligned");
    TcProblem pr;
    pr.w_per_sample = 0;
    pr.src = x; pr.SN = g->N; pr.SH = g->H; pr.SW = g->W; pr.SC = g->C;
    pr.wmat = w; pr.Ncol = g->K; pr.Ktot = g->R * g->S * g->C;
    pr.out = y; pr.OH = g->P; pr.OW = g->Q;
    pr.o_mul = 1; pr.o_offy = 0; pr.o_offx = 0; pr.FH = g->P; pr.FW = g->Q;
    pr.stride = g->stride; pr.ntaps = g->R * g->S;
    for (int r = 0; r < g->R; ++r)
        for (int s = 0; s < g->S; ++s) {
            const int t = r * g->S + s;
            pr.oy[t] = r - g->pad_t; pr.ox[t] = s - g->pad_l; pr.wk[t] = t * g->C;
        }
    return tc_dispatch_split(pr, e, st);
}

// Per-sample filters (the style-modulated convolution, stylegan2_layers.py:284-323): image n is convolved with filter matrix n.
// Only the shared-window kernel implements it, for stride-1 problems whose 16 x 8 tiles cover the map exactly with an even
// number of tiles per image (both CTAs of a pair then read the same filter); anything else reports UNSUPPORTED and the
// caller scales the input instead.
bool tc_per_sample_eligible(const sae_conv_geom* g, int dgrad) {
    if (g->stride != 1 || g->R * g->S > TC_MAX_TAPS) return false;
    const int src_c = dgrad ? g->K : g->C, ncol = dgrad ? g->C : g->K;
    const int oh = dgrad ? g->H : g->P, ow = dgrad ? g->W : g->Q;
    if (src_c % 32 != 0 || ncol % 32 != 0 || oh % 16 != 0 || ow % 8 != 0) return false;
    if (((oh / 16) * (ow / 8)) % 2 != 0) return false;
    if (g->R > 3 || g->S > 3) return false;
    return (int64_t)g->N * oh * ow >= 2 * 128;
}

int tc_conv_per_sample(const float* src, const float* w, float* out, const sae_conv_geom* g, int dgrad, const EpiParams& e, cudaStream_t st) {
    if (!tc_per_sample_eligible(g, dgrad)) return fail(SAE_E_UNSUPPORTED, "per-sample conv: shape outside the shared-window kernel");
    if (!ptr_ok(src, w, out)) return fail(SAE_E_INVALID, "per-sample conv: pointers must be 16-byte aligned");
    TcProblem pr;
    pr.w_per_sample = 1;
    pr.stride = 1; pr.o_mul = 1; pr.o_offy = 0; pr.o_offx = 0; pr.ntaps = g->R * g->S;
    pr.src = src; pr.wmat = w; pr.out = out; pr.SN = g->N;
    if (!dgrad) {
        pr.SH = g->H; pr.SW = g->W; pr.SC = g->C; pr.Ncol = g->K; pr.Ktot = g->R * g->S * g->C;
        pr.OH = g->P; pr.OW = g->Q; pr.FH = g->P; pr.FW = g->Q;
    } else {
        pr.SH = g->P; pr.SW = g->Q; pr.SC = g->K; pr.Ncol = g->C; pr.Ktot = g->R * g->S * g->K;
        pr.OH = g->H; pr.OW = g->W; pr.FH = g->H; pr.FW = g->W;
    }
    for (int r = 0; r < g->R; ++r)
        for (int s_ = 0; s_ < g->S; ++s_) {
            const int t = r * g->S + s_;
            pr.oy[t] = dgrad ? g->pad_t - r : r - g->pad_t;
            pr.ox[t] = dgrad ? g->pad_l - s_ : s_ - g->pad_l;
            pr.wk[t] = t * pr.SC;
        }
    if (!window_ok(pr)) return fail(SAE_E_UNSUPPORTED, "per-sample conv: tap offsets outside the shared window");
    if (pr.Ncol % 128 == 0) return tc5_launch<128>(pr, e, st);
    return pr.Ncol % 64 == 0 ? tc5_launch<64>(pr, e, st) : tc5_launch<32>(pr, e, st);
}

// One conv_tc5m launch over the rectangle all four parity classes share, then each class's thin remainder strips through
// the ordinary per-class path.  Returns SAE_E_UNSUPPORTED (quietly) when the shape is outside the merged kernel's reach.
// ACCS = 1: one class per work item; ACCS = 2: two classes per work item (see the kernel's header).
template <int ACCS>
static int tc_dgrad_merged(const TcProblem* cp, const EpiParams& e, cudaStream_t st) {
    const TcProblem& p0 = cp[0];
    if (p0.Ncol % 128 != 0 || p0.SC % 32 != 0) return SAE_E_UNSUPPORTED;
    int MH = cp[0].OH, MW = cp[0].OW;
    int oy_min = cp[0].oy[0], oy_max = oy_min, ox_min = cp[0].ox[0], ox_max = ox_min;
    for (int c = 0; c < TC_MAX_CLS; ++c) {
        if (cp[c].ntaps < 1 || cp[c].ntaps > 4) return SAE_E_UNSUPPORTED;
        MH = cp[c].OH < MH ? cp[c].OH : MH; MW = cp[c].OW < MW ? cp[c].OW : MW;
        for (int t = 0; t < cp[c].ntaps; ++t) {
            oy_min = cp[c].oy[t] < oy_min ? cp[c].oy[t] : oy_min; oy_max = cp[c].oy[t] > oy_max ? cp[c].oy[t] : oy_max;
            ox_min = cp[c].ox[t] < ox_min ? cp[c].ox[t] : ox_min; ox_max = cp[c].ox[t] > ox_max ? cp[c].ox[t] : ox_max;
        }
    }
    if (MH < 16 || MW < 8 || oy_max - oy_min > 2 || ox_max - ox_min > 2) return SAE_E_UNSUPPORTED;
    if ((int64_t)p0.SN * MH * MW < 2 * 128) return SAE_E_UNSUPPORTED;
    if ((16 + oy_max - oy_min) * (8 + ox_max - ox_min) * 128 > tc5m_aslot<ACCS>()) return SAE_E_UNSUPPORTED;      // window slot

    TcProblem main = p0;                       // geometry of the shared rectangle (taps / offsets come from the group tables)
    main.OH = MH; main.OW = MW;
    TcParams p;
    tc_fill_params(main, e, p);
    p.tw = 8; p.th = 16; p.tn = 1;
    p.tiles_w = (MW + p.tw - 1) / p.tw;
    p.tiles_h = (MH + p.th - 1) / p.th;
    p.tiles_n = main.SN;
    p.oy_min = oy_min; p.ox_min = ox_min;
    p.ww = p.tw + (ox_max - ox_min);
    p.wh = p.th + (oy_max - oy_min);
    TcGroups grp = {};
    TcOutMaps outs;
    CUtensorMap msrc, mw, mdummy;
    int rc = tc_encode_maps(main, p, 128 / 2, &msrc, &mw, &mdummy);      // weight map (the others are rebuilt below)
    if (rc) return rc;
    {
        cuuint64_t dims[4] = {(cuuint64_t)main.SC, (cuuint64_t)main.SW, (cuuint64_t)main.SH, (cuuint64_t)main.SN};
        cuuint64_t strides[3] = {(cuuint64_t)main.SC * 4, (cuuint64_t)main.SW * main.SC * 4, (cuuint64_t)main.SH * main.SW * main.SC * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.ww, (cuuint32_t)p.wh, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        rc = encode_map(&msrc, main.src, 4, dims, strides, box, es);
        if (rc) return rc;
    }
    for (int c = 0; c < TC_MAX_CLS; ++c) {
        const TcProblem& q = cp[c];
        grp.o_offy[c] = q.o_offy; grp.o_offx[c] = q.o_offx;
        // output sub-grid of class c restricted to the shared MH x MW rectangle (the tensor map's extent clips the stores)
        cuuint64_t dims[4] = {(cuuint64_t)q.Ncol, (cuuint64_t)MW, (cuuint64_t)MH, (cuuint64_t)q.SN};
        cuuint64_t strides[3] = {(cuuint64_t)q.o_mul * q.Ncol * 4, (cuuint64_t)q.o_mul * q.FW * q.Ncol * 4,
                                 (cuuint64_t)q.FH * q.FW * q.Ncol * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.tw, (cuuint32_t)p.th, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        rc = encode_map(&outs.m[c], q.out + ((int64_t)q.o_offy * q.FW + q.o_offx) * q.Ncol, 4, dims, strides, box, es);
        if (rc) return rc;
    }
    // groups: classes sorted by tap count; ACCS = 2 pairs the longest with the shortest (5 + 4 taps for a 3x3 filter)
    int order[TC_MAX_CLS] = {0, 1, 2, 3};
    for (int i = 0; i < TC_MAX_CLS; ++i)
        for (int j = i + 1; j < TC_MAX_CLS; ++j)
            if (cp[order[j]].ntaps > cp[order[i]].ntaps) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
    grp.ngrp = TC_MAX_CLS / ACCS;
    for (int g = 0; g < grp.ngrp; ++g) {
        int nt = 0;
        for (int a = 0; a < ACCS; ++a) {
            const int c = ACCS == 1 ? g : (a == 0 ? order[g] : order[TC_MAX_CLS - 1 - g]);
            grp.cls[g][a] = c;
            const TcProblem& q = cp[c];
            for (int t = 0; t < q.ntaps; ++t, ++nt) {
                grp.wk[g][nt] = q.wk[t];
                grp.arow[g][nt] = (unsigned short)((q.oy[t] - oy_min) * p.ww + (q.ox[t] - ox_min));
                grp.acc[g][nt] = (unsigned char)a;
                grp.first[g][nt] = (unsigned char)(t == 0);
            }
        }
        grp.ntaps[g] = nt;
    }
    constexpr size_t smem = tc5m_smem_bytes<ACCS>();
    static bool attr_done = false;
    if (!attr_done) {
        SAE_CUDA_TRY(cudaFuncSetAttribute(conv_tc5m_kernel<ACCS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const int tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int n_blocks = main.Ncol / 128;
    const int total_work = ((tiles + 1) / 2) * grp.ngrp * n_blocks;
    int clusters = ACCS == 1 ? sm_count() : sm_count() / 2;      // two CTAs per SM / one CTA per SM, 2 CTAs per cluster
    if (clusters > total_work) clusters = total_work;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * 2), 1, 1);
    cfg.blockDim = dim3(64 + 128 * ACCS, 1, 1);      // producer warp, MMA warp, four epilogue warps per class accumulator
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SAE_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_tc5m_kernel<ACCS>, msrc, mw, outs, p, grp, n_blocks, total_work));
    rc = check_launch("conv_tc5m");
    if (rc) return rc;
    // remainders of each class beyond the shared rectangle: right strip (full height, takes the corner), bottom strip —
    // all of them in ONE launch of the one-tile kernel (profiles/r2_wgrad_pairs.txt: 0.597 ms with four strip launches
    // against 0.523 ms for the strip-free 256^2 shape)
    TcProblem strips[2 * TC_MAX_CLS];
    int ns = 0;
    for (int c = 0; c < TC_MAX_CLS; ++c) {
        const TcProblem& q = cp[c];
        if (MW < q.OW) strips[ns++] = tc_subproblem(q, 0, q.OH, MW, q.OW);
        if (MH < q.OH) strips[ns++] = tc_subproblem(q, MH, q.OH, 0, MW);
    }
    for (int i = 0; i < ns; i += TC_MAX_BATCH) {
        const int cnt = ns - i < TC_MAX_BATCH ? ns - i : TC_MAX_BATCH;
        rc = tc_launch_batch<128>(strips + i, cnt, e, st);
        if (rc) return rc;
    }
    return SAE_OK;
}

int tc_dgrad(const float* dy, const float* wt, float* dx, const sae_conv_geom* g, const EpiParams& e, cudaStream_t st) {
    if (!ptr_ok(dy, wt, dx)) return fail(SAE_E_INVALID, "conv2d_dgrad(tcgen05): pointers must be 16-byte aligned");
    TcProblem pr;
    pr.w_per_sample = 0;
    pr.src = dy; pr.SN = g->N; pr.SH = g->P; pr.SW = g->Q; pr.SC = g->K;
    pr.wmat = wt; pr.Ncol = g->C; pr.Ktot = g->R * g->S * g->K;
    pr.stride = 1; pr.FH = g->H; pr.FW = g->W;
    const int st_ = g->stride;
    if (st_ == 1) {
        pr.out = dx; pr.OH = g->H; pr.OW = g->W; pr.o_mul = 1; pr.o_offy = 0; pr.o_offx = 0;
        pr.ntaps = g->R * g->S;
        for (int r = 0; r < g->R; ++r)
            for (int s = 0; s < g->S; ++s) {
                const int t = r * g->S + s;
                pr.oy[t] = g->pad_t - r; pr.ox[t] = g->pad_l - s; pr.wk[t] = t * g->K;
            }
        return tc_dispatch_split(pr, e, st);
    }
    // stride 2 (the generator's transposed convolution and the data-gradient of the strided convs): the output
    // splits into 4 parity classes (ho, wo); class outputs x[2i+ho, 2j+wo] only see taps with r = (ho + pad_t) mod 2,
    // s = (wo + pad_l) mod 2, read at source offset (ho + pad_t - r) / 2 — four dense stride-1 problems writing
    // interleaved sub-grids (the output tensor map carries the doubled strides).
    bool need_zero = false;
    for (int ho = 0; ho < 2; ++ho)
        for (int wo = 0; wo < 2; ++wo) {
            const int rp = (ho + g->pad_t) & 1, sp = (wo + g->pad_l) & 1;
            if (rp >= g->R || sp >= g->S) need_zero = true;
        }
    if (need_zero) SAE_CUDA_TRY(cudaMemsetAsync(dx, 0, (size_t)g->N * g->H * g->W * g->C * sizeof(float), st));
    TcProblem cls_pr[TC_MAX_CLS];
    int ncls = 0;
    for (int ho = 0; ho < 2; ++ho)
        for (int wo = 0; wo < 2; ++wo) {
            const int rp = (ho + g->pad_t) & 1, sp = (wo + g->pad_l) & 1;
            if (rp >= g->R || sp >= g->S) continue;
            if (ho >= g->H || wo >= g->W) continue;
            pr.out = dx; pr.o_mul = 2; pr.o_offy = ho; pr.o_offx = wo;
            pr.OH = (g->H - ho + 1) / 2; pr.OW = (g->W - wo + 1) / 2;
            int nt = 0;
            for (int r = rp; r < g->R; r += 2)
                for (int s = sp; s < g->S; s += 2) {
                    // exact division: (ho + pad_t - r) is even; C++ division truncates toward zero, so floor by hand
                    const int ny = ho + g->pad_t - r, nx = wo + g->pad_l - s;
                    pr.oy[nt] = ny >= 0 ? ny / 2 : -((-ny) / 2);
                    pr.ox[nt] = nx >= 0 ? nx / 2 : -((-nx) / 2);
                    pr.wk[nt] = (r * g->S + s) * g->K;
                    ++nt;
                }
            pr.ntaps = nt;
            cls_pr[ncls++] = pr;
        }
    static int merged = -1;
    // SAE_DGRAD_MERGED: 2 (default) = conv_tc5m with two classes per work item; 1 = one class per work item (386 -> 416,
    // 441 -> 506 TFLOP/s on the discriminator shapes over 0, profiles/r2_conv_bench_s2.txt); 0 = one launch per class (A/B runs)
    if (merged < 0) { const char* v = getenv("SAE_DGRAD_MERGED"); merged = v ? atoi(v) : 2; }
    if (merged && ncls == TC_MAX_CLS && !need_zero) {
        int rc = merged >= 2 ? tc_dgrad_merged<2>(cls_pr, e, st) : SAE_E_UNSUPPORTED;
        if (rc == SAE_E_UNSUPPORTED) rc = tc_dgrad_merged<1>(cls_pr, e, st);
        if (rc != SAE_E_UNSUPPORTED) return rc;
    }
    for (int c = 0; c < ncls; ++c) {
        int rc = tc_dispatch_split(cls_pr[c], e, st);
        if (rc) return rc;
    }
    return SAE_OK;
}

}  // namespace sae
