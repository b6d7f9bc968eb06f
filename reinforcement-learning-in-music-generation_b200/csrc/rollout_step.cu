// The recurrent token step of the rollout as ONE persistent cooperative kernel  (cpm_rollout_create / cpm_rollout_run).
// OPT-IN (RolloutEngine(mode="persistent") / CPM_ROLLOUT_MODE=persistent): measured on B200 at 256 songs it runs a token step in
// 700-760 us against 442 us for the default chain of 67 kernels (profiles/r02_summary.md, "persistent rollout step"); it is kept as
// the one non-default rollout mode because it is complete, parity-tested, and the per-stage timeline it records
// (cpm_debug_rollout_timing, tools/phase_timing_rollout.py) is the measurement that explains where a token step's time goes.
//
// What it replaces: the reference generates one token of one song per host round trip (testing-no-type-cp.py:157-167:
// forward_hidden(is_training=False) -> forward_output -> six numpy samplers).  Batched over 256 songs the step is a chain of
// 63 DEPENDENT stages (embedding, 12 x [QKV, state, out-projection, linear1, linear2], heads, sample) with 0.3 us of math each,
// against an HBM floor of 139 us (77 MB of bf16 weights + 830 MB of recurrent-state traffic).
//
// Design (one CTA per SM, 320 threads, 225 KB of shared memory, 32 TMEM columns):
//   * warp 0 = WEIGHT PRODUCER.  The weights do not depend on the chain, so one thread streams the weight tiles of THIS CTA's
//     output tiles, for every stage of every token, through a 9 x 16 KB TMA ring - it never waits for a device-wide barrier and
//     runs ahead of the compute by the depth of the ring (the next stage's weights are in shared memory when its barrier opens).
//   * warp 1 = UMMA ISSUER.  Linear layers run "swap-AB": the UMMA M axis carries 64 / 128 weight rows (output features), the N
//     axis 16 / 32 songs, so a tile needs BR x K activations (16-64 KB) and the accumulator is [features x songs] in TMEM.
//   * warps 2-9 = WORKERS (256 threads): stage the activation tile (L2 -> registers -> swizzled shared memory, applying the
//     LayerNorm(s) that precede the Linear on the way - or gathering the CP embeddings for in_linear), run the epilogue
//     (bias, GELU, positional encoding), the recurrent-state stage (each warp streams whole (song, head) tiles through private
//     bulk-copy slots, summation order of linattn_step_kernel) and the sampler (heads_dev.cuh, the code of cpm_heads_sample).
//   * stages are separated by a device-wide barrier among the workers (one 64-bit ticket counter in L2); activations that cross
//     CTAs are written with plain stores + a release and read with ld.global.cg (L2), never through L1.
//   * LayerNorm placement: out-projection and linear2 write their plain outputs o (bf16); the stage that consumes them forms
//     x + o in fp32 and normalises its rows while staging them (one warp per 2-4 rows, the arithmetic of ln_residual_fwd_kernel)
//     and the tiles with weight-tile index 0 also write the normalised rows out, because they are the next residual input.
//     Every value is rounded to bf16 exactly where the kernel chain rounds it; what still differs is the fp32 accumulation
//     inside the tensor core (weights on M here, songs on M in cpm_gemm_nt_small): about one bf16 ulp in one of ~10^5 outputs.
//
// Why it is not faster (stage timeline at 256 songs, us): device-wide barrier 1.4; activation tile landed 1.4 after the barrier;
// LayerNorm of 4 rows per warp 2.8 (8 warps per SM is all the thread-level parallelism a stage has); 32 UMMAs 1.5-2.3 (a
// small-N UMMA costs 46 cycles whatever its N - tools/probes/umma_small_n.cu - plus one mbarrier round trip per ring stage);
// epilogue 1-2.4; state stage 14 (HBM-bound, 138 MB per layer).  5-10 us per stage x 63 stages; the kernel chain pays 5.6 us per
// kernel with far more warps in flight per stage and programmatic dependent launch hiding each prologue.
#include "cpm_common.cuh"
#include "tc_common.cuh"
#include "heads_dev.cuh"
#include <new>

namespace cpm {
namespace {
using namespace tc;

constexpr int RS_THREADS = 320, RS_WORKERS = 256, RS_NS = 9, RS_MAX_PH = 80;
constexpr uint32_t RS_STAGE = 16384, RS_A_BYTES = 65536;
enum { PH_GEMM = 0, PH_STATE = 1, PH_SAMPLE = 2 };
enum { PRO_PLAIN = 0, PRO_LN = 1, PRO_LN2 = 2, PRO_EMBED = 3 };
enum { EPI_STORE = 0, EPI_GELU = 1, EPI_PE = 3 };

struct RsPhase {
    int type, N, K, BW, BR, n_wtiles, n_rtiles, pro, epi, lda, ldd, ldr, pad_;
    const __nv_bfloat16 *A;
    const float *g1, *b1, *g2, *b2;
    __nv_bfloat16 *xout;
    const float *bias;
    __nv_bfloat16 *D;
    const __nv_bfloat16 *R;               // LayerNorm prologue: the residual stream added to A before normalising
    float *S, *Z;
    void *pad2_[2];
};

struct RsEmbed {
    const float *tables[CPM_MAX_ATTR];
    int n_tokens[CPM_MAX_ATTR], emb[CPM_MAX_ATTR], off[CPM_MAX_ATTR + 1];
    float scale[CPM_MAX_ATTR];
    int n_attr;
};

struct __align__(256) RsPlan {
    CUtensorMap tm[RS_MAX_PH];            // weight tensor map of GEMM phase i (64-byte aligned entries)
    RsPhase ph[RS_MAX_PH];
    int n_phases, B, H, d_model, logits_ld, pe_len, true_positions, greedy, max_steps, pad_;
    float ln_eps, attn_eps;
    uint64_t seed;
    int64_t seq_base;
    RsEmbed emb;
    SegParams seg;
    const float *pe;
    int64_t *cur;
    float *logp;
    int64_t *hist_tok;
    float *hist_logp;
    int32_t *step_dev;
    const __nv_bfloat16 *qkv;
    __nv_bfloat16 *attn;
    const __nv_bfloat16 *logits;
    unsigned long long *barrier;
    int *err_flag;
};

constexpr uint32_t RS_OFF_A = RS_NS * RS_STAGE;
constexpr uint32_t RS_OFF_PH = RS_OFF_A + RS_A_BYTES;
constexpr uint32_t RS_OFF_BAR = RS_OFF_PH + RS_MAX_PH * (uint32_t)sizeof(RsPhase);
constexpr uint32_t RS_SMEM = RS_OFF_BAR + 512;                             // w_full[9] w_empty[9] a_full acc_full s_full[8][3] + TMEM slot
static_assert(RS_SMEM <= 232448, "rollout step: shared memory");
static_assert(sizeof(RsPhase) % 16 == 0, "RsPhase is copied with 16-byte accesses");

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldcg16(const void *p) { return __ldcg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ float ldcg_bf16(const __nv_bfloat16 *p) {
    const unsigned short u = __ldcg(reinterpret_cast<const unsigned short *>(p));
    return __uint_as_float((uint32_t)u << 16);
}
__device__ __forceinline__ void unpack8(const uint4 &r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 16-byte chunk `c8` (8 bf16) of row `row` of a [BR rows x K] K-major SWIZZLE_128B operand: k-block kb = c8 / 8 is a [BR x 128 B] tile
__device__ __forceinline__ uint32_t a_chunk_off(int BR, int row, int c8) {
    return (uint32_t)(c8 >> 3) * (uint32_t)(BR * 128) + sw128_off(row, c8 & 7);
}

// Device-wide barrier among the worker threads of all CTAs.  Monotonic 64-bit ticket counter: round r completes when the
// counter reaches (r + 1) * gridDim.x.  Writers: plain stores, bar.sync, then the leader's red.release.gpu (cumulative over what
// the bar.sync ordered before it); readers: the leader's relaxed polls + one acquire fence, bar.sync, then L2 loads (ld.global.cg, TMA).
// The spin is bounded: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned long long *ctr, unsigned long long &target, int wtid) {
    worker_sync();
    if (wtid == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
        target += gridDim.x;
        uint32_t spins = 0;
        while (ld_relaxed_u64(ctr) < target) {                    // relaxed polls: an acquire load would flush L1 on every iteration
            if (++spins > (1u << 26)) { printf("cpmusic: rollout-step grid barrier timed out (block %d)\n", (int)blockIdx.x); __trap(); }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    worker_sync();
}

// ---------------------------------------------------------------- activation staging (workers)
__device__ __forceinline__ void stage_plain(uint8_t *sa, const RsPhase &p, int row0, int B, int wtid) {
    const int KC = p.K >> 3, KC8 = ((p.K + 63) >> 6) << 3, total = p.BR * KC8;
    for (int c0 = wtid; c0 < total; c0 += 8 * RS_WORKERS) {          // 8 x 16 B per thread in flight
        uint4 v[8];
        int rw[8], c8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * RS_WORKERS;
            rw[u] = c / KC8;
            c8[u] = c - rw[u] * KC8;
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            if (c < total && row0 + rw[u] < B && c8[u] < KC) v[u] = ldcg16(p.A + (int64_t)(row0 + rw[u]) * p.lda + c8[u] * 8);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (c0 + u * RS_WORKERS < total) *reinterpret_cast<uint4 *>(sa + a_chunk_off(p.BR, rw[u], c8[u])) = v[u];
    }
}

__device__ __forceinline__ void stage_embed(uint8_t *sa, const RsPhase &p, const RsPlan *plan, int row0, int B, int wtid) {
    const RsEmbed &e = plan->emb;
    const int KC = p.K >> 3, KC8 = ((p.K + 63) >> 6) << 3, total = p.BR * KC8;
    for (int c0 = wtid; c0 < total; c0 += 5 * RS_WORKERS) {          // 5 chunks per thread per pass: ids, then tables, then stores
        int rw[5], c8[5], at[5];
        int64_t id[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int c = c0 + u * RS_WORKERS;
            rw[u] = c / KC8;
            c8[u] = c - rw[u] * KC8;
            id[u] = -1;
            at[u] = 0;
            if (c < total && row0 + rw[u] < B && c8[u] < KC) {
                int a = 0;
                while (c8[u] * 8 >= e.off[a + 1]) ++a;
                at[u] = a;
                id[u] = __ldcg(plan->cur + (int64_t)(row0 + rw[u]) * e.n_attr + a);
            }
        }
        float4 x[5], y[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            y[u] = x[u];
            const bool live = c0 + u * RS_WORKERS < total && row0 + rw[u] < B && c8[u] < KC;
            if (live && (id[u] < 0 || id[u] >= e.n_tokens[at[u]])) {
                if (plan->err_flag) atomicExch(plan->err_flag, 1);
            } else if (live) {
                const float *src = e.tables[at[u]] + id[u] * e.emb[at[u]] + (c8[u] * 8 - e.off[at[u]]);
                x[u] = __ldg(reinterpret_cast<const float4 *>(src));
                y[u] = __ldg(reinterpret_cast<const float4 *>(src + 4));
            }
        }
#pragma unroll
        for (int u = 0; u < 5; ++u)
            if (c0 + u * RS_WORKERS < total) {
                const float sc = e.scale[at[u]];
                const float o[8] = {x[u].x * sc, x[u].y * sc, x[u].z * sc, x[u].w * sc, y[u].x * sc, y[u].y * sc, y[u].z * sc, y[u].w * sc};
                *reinterpret_cast<uint4 *>(sa + a_chunk_off(p.BR, rw[u], c8[u])) = pack8(o);
            }
    }
}

// LayerNorm of NR rows held by a warp (16 values per lane and row: chunks `lane` and `lane + 32` of 8), in place; per row the
// arithmetic of ln_residual_fwd_kernel: fp32 two-pass statistics, (v - mean) * rstd * gamma + beta.  The NR butterflies run
// side by side (NR independent shuffles in flight per level).  gm / bt: this lane's 16 gammas / betas.
template <int NR>
__device__ __forceinline__ void ln_rows(float (&v)[NR][16], int G, int lane, int d, float eps, const float (&gm)[16], const float (&bt)[16]) {
    float s[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        s[r] = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (lane + 32 * i < G) {
#pragma unroll
                for (int j = 0; j < 8; ++j) s[r] += v[r][8 * i + j];
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < NR; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    float mean[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        mean[r] = s[r] / (float)d;
        s[r] = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (lane + 32 * i < G) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float c = v[r][8 * i + j] - mean[r]; s[r] += c * c; }
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < NR; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const float rstd = rsqrtf(s[r] / (float)d + eps);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[r][j] = (v[r][j] - mean[r]) * rstd * gm[j] + bt[j];
    }
}
__device__ __forceinline__ void load_affine(float (&gm)[16], float (&bt)[16], const float *__restrict__ gamma, const float *__restrict__ beta, int G,
                                            int lane) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int g = lane + 32 * i;
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, b0 = g0, b1 = g0;
        if (g < G) {
            g0 = __ldg(reinterpret_cast<const float4 *>(gamma + g * 8)); g1 = __ldg(reinterpret_cast<const float4 *>(gamma + g * 8 + 4));
            b0 = __ldg(reinterpret_cast<const float4 *>(beta + g * 8)); b1 = __ldg(reinterpret_cast<const float4 *>(beta + g * 8 + 4));
        }
        gm[8 * i + 0] = g0.x; gm[8 * i + 1] = g0.y; gm[8 * i + 2] = g0.z; gm[8 * i + 3] = g0.w;
        gm[8 * i + 4] = g1.x; gm[8 * i + 5] = g1.y; gm[8 * i + 6] = g1.z; gm[8 * i + 7] = g1.w;
        bt[8 * i + 0] = b0.x; bt[8 * i + 1] = b0.y; bt[8 * i + 2] = b0.z; bt[8 * i + 3] = b0.w;
        bt[8 * i + 4] = b1.x; bt[8 * i + 5] = b1.y; bt[8 * i + 6] = b1.z; bt[8 * i + 7] = b1.w;
    }
}

// rows of the tile normalised on the way into shared memory (K = d_model <= 512): warp w stages rows w, w + 8, ...;
// row = LayerNorm(R + A) with R the residual stream and A the preceding Linear's output (both bf16, summed in fp32).
// Every global load of the warp (its 2 or 4 rows of both sources, gamma and beta) is in flight before the first use.
template <int NR>
__device__ __forceinline__ void stage_ln_rows(uint8_t *sa, const RsPhase &p, int row0, int B, int ww, int lane, float eps, bool write_x,
                                              unsigned long long *stamp) {
    const int G = p.K >> 3, KC8 = ((p.K + 63) >> 6) << 3;
    uint4 ra[NR][2], rx[NR][2];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ra[r][i] = make_uint4(0u, 0u, 0u, 0u);
            rx[r][i] = make_uint4(0u, 0u, 0u, 0u);
            const int row = row0 + ww + 8 * r;
            if (row < B && lane + 32 * i < G) {
                ra[r][i] = ldcg16(p.A + (int64_t)row * p.lda + (lane + 32 * i) * 8);
                rx[r][i] = ldcg16(p.R + (int64_t)row * p.ldr + (lane + 32 * i) * 8);
            }
        }
    float gm[16], bt[16];
    load_affine(gm, bt, p.g1, p.b1, G, lane);
    if (stamp) {                                                  // development aid: when did the loads land
        if (ra[NR - 1][1].x == 0x7fc07fc1u && rx[NR - 1][1].y == 0x7fc07fc1u && gm[15] == 123.f) stamp[4] = 1;
        stamp[4] = globaltimer_ns();
    }
    float v[NR][16];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float ta[8], tx[8];
            unpack8(ra[r][i], ta);
            unpack8(rx[r][i], tx);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[r][8 * i + j] = tx[j] + ta[j];
        }
    ln_rows<NR>(v, G, lane, p.K, eps, gm, bt);
    if (p.pro == PRO_LN2) {                                      // the layer's norm2, then the encoder's final norm (bf16 in between)
        load_affine(gm, bt, p.g2, p.b2, G, lane);
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int j = 0; j < 16; ++j) v[r][j] = bf16_round(v[r][j]);
        ln_rows<NR>(v, G, lane, p.K, eps, gm, bt);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int rw = ww + 8 * r, row = row0 + rw;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int g = lane + 32 * i;
            if (g < KC8) {
                float t[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) t[j] = (g < G && row < B) ? v[r][8 * i + j] : 0.f;
                const uint4 o = pack8(t);
                *reinterpret_cast<uint4 *>(sa + a_chunk_off(p.BR, rw, g)) = o;
                if (write_x && g < G && row < B) *reinterpret_cast<uint4 *>(p.xout + (int64_t)row * p.K + g * 8) = o;
            }
        }
    }
}
__device__ __forceinline__ void stage_ln(uint8_t *sa, const RsPhase &p, int row0, int B, int ww, int lane, float eps, bool write_x,
                                         unsigned long long *stamp) {
    if (p.BR == 32) stage_ln_rows<4>(sa, p, row0, B, ww, lane, eps, write_x, stamp);
    else stage_ln_rows<2>(sa, p, row0, B, ww, lane, eps, write_x, stamp);
    if (stamp) stamp[5] = globaltimer_ns();
}

// ---------------------------------------------------------------- epilogue (workers): TMEM [features x songs] -> global
template <int NC>
__device__ __forceinline__ void epilogue_cols(const uint32_t (&r)[NC], const RsPhase &p, int n, bool n_ok, int row0, int col0, int B, float bias,
                                              float pe) {
    if (!n_ok) return;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const int row = row0 + col0 + j;
        if (row < B) {
            float x = __uint_as_float(r[j]) + bias;
            if (p.epi == EPI_GELU) x = gelu_f<false>(bf16_round(x));
            else if (p.epi == EPI_PE) x = bf16_round(x) + pe;
            p.D[(int64_t)row * p.ldd + n] = __float2bfloat16_rn(x);
        }
    }
}

// ---------------------------------------------------------------- recurrent state stage (workers)
// Each WARP owns whole (song, head) tiles and never synchronises with the other warps: the 64 x 64 fp32 state streams through
// three warp-private 2 KB shared-memory slots (cp.async.bulk, 8 rows per copy, up to 6 KB in flight per warp and 48 KB per SM -
// the bandwidth-delay product of HBM at 148 SMs), lane l owns columns 2l, 2l + 1 of every row, updates S in place
// (S_em = fma(phi(k)_e, v_m, S_em), written straight back from registers) and accumulates out_m = sum_e phi(q)_e S_em in the
// summation order of linattn_step_kernel: groups of 8 rows combined as ((0+1)+(2+3))+((4+5)+(6+7)), groups added in order.
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void st_stream2(float *p, float a, float b) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float tree8(const float (&a)[8]) { return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7])); }

__device__ __forceinline__ void state_phase(uint8_t *sa, uint64_t *s_full, uint32_t &scnt, const RsPhase &p, const RsPlan *plan, int cta, int G, int ww,
                                            int lane) {
    const int H = plan->H, dm = plan->d_model, tiles = plan->B * H;
    uint8_t *slots = sa + ww * 8192;
    float *kq = reinterpret_cast<float *>(slots + 6144), *qz = kq + 128;      // (phi(k)_e, phi(q)_e) pairs; phi(q)_e * z_e
    uint64_t *bars = s_full + ww * 3;
    int n_my = 0;
    for (int i = ww; cta + G * i < tiles; i += 8) ++n_my;
    const int total = n_my * 8;                                               // 8-row groups this warp streams
    auto issue = [&](int gi, uint32_t cnt) {                                  // lane 0: group gi -> slot cnt % 3
        const int tile = cta + G * (ww + 8 * (gi >> 3)), sl = cnt % 3;
        mbar_expect_tx(bars + sl, 2048u);
        bulk_load(slots + sl * 2048, p.S + (int64_t)tile * 4096 + (gi & 7) * 512, 2048u, bars + sl);
    };
    if (lane == 0)
        for (int gi = 0; gi < 3 && gi < total; ++gi) issue(gi, scnt + gi);
    float acc0 = 0.f, acc1 = 0.f, den = 1.f, v0 = 0.f, v1 = 0.f;
    int tile = 0, n = 0, h = 0;
    for (int gi = 0; gi < total; ++gi, ++scnt) {
        const int g = gi & 7;
        if (g == 0) {                                                         // ---- per-tile set-up: q, k, v, normaliser
            tile = cta + G * (ww + 8 * (gi >> 3));
            n = tile / H;
            h = tile - n * H;
            const __nv_bfloat16 *q = plan->qkv + (int64_t)n * 3 * dm + h * 64;
            const float q0 = ldcg_bf16(q + lane), q1 = ldcg_bf16(q + lane + 32), k0 = ldcg_bf16(q + dm + lane), k1 = ldcg_bf16(q + dm + lane + 32);
            const uint32_t vv = __ldcg(reinterpret_cast<const unsigned int *>(q + 2 * dm + 2 * lane));
            float *z = p.Z + (int64_t)tile * 64;
            const float z0 = z[lane], z1 = z[lane + 32];
            const float kf0 = phi(k0), kf1 = phi(k1), qf0 = phi(q0), qf1 = phi(q1);
            const float zn0 = z0 + kf0, zn1 = z1 + kf1;
            z[lane] = zn0;
            z[lane + 32] = zn1;
            v0 = __uint_as_float(vv << 16);
            v1 = __uint_as_float(vv & 0xFFFF0000u);
            __syncwarp();                                                     // the previous tile's readers of kq / qz are done
            *reinterpret_cast<float2 *>(kq + 2 * lane) = make_float2(kf0, qf0);
            *reinterpret_cast<float2 *>(kq + 2 * (lane + 32)) = make_float2(kf1, qf1);
            qz[lane] = qf0 * zn0;
            qz[lane + 32] = qf1 * zn1;
            __syncwarp();
            float dp = 0.f;
            if (lane < 8) {
                float a[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = qz[8 * lane + i];
                dp = tree8(a);
            }
            den = plan->attn_eps;
#pragma unroll
            for (int w = 0; w < 8; ++w) den += __shfl_sync(0xffffffffu, dp, w);
            acc0 = 0.f;
            acc1 = 0.f;
        }
        const int sl = scnt % 3;
        mbar_wait(bars + sl, (scnt / 3) & 1);
        const uint8_t *src = slots + sl * 2048 + lane * 8;
        float *dst = p.S + (int64_t)tile * 4096 + g * 512 + 2 * lane;
        float a0[8], a1[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float2 sv = *reinterpret_cast<const float2 *>(src + r * 256);
            const float2 kqe = *reinterpret_cast<const float2 *>(kq + 2 * (8 * g + r));
            const float s0 = fmaf(kqe.x, v0, sv.x), s1 = fmaf(kqe.x, v1, sv.y);
            st_stream2(dst + r * 64, s0, s1);
            a0[r] = kqe.y * s0;
            a1[r] = kqe.y * s1;
        }
        acc0 += tree8(a0);
        acc1 += tree8(a1);
        __syncwarp();                                                         // every lane has consumed the slot
        if (lane == 0 && gi + 3 < total) issue(gi + 3, scnt + 3);
        if (g == 7) *reinterpret_cast<uint32_t *>(plan->attn + (int64_t)n * dm + h * 64 + 2 * lane) = pack_bf16(acc0 / den, acc1 / den);
    }
    asm volatile("fence.proxy.async;" ::: "memory");        // next token: these generic-proxy stores are read back by the bulk-copy engine
}

// ---------------------------------------------------------------- the kernel

// `timing` (development aid, NULL by default): per CTA 8 stamp slots per stage of the first RS_TIMED_STEPS steps - work done, barrier passed, activation tile staged, accumulator ready
constexpr int RS_TIMED_STEPS = 4;
__global__ void __launch_bounds__(RS_THREADS, 1) rollout_step_kernel(const RsPlan *__restrict__ plan, int n_steps, unsigned long long *timing) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t *sa = sm + RS_OFF_A;
    RsPhase *phs = reinterpret_cast<RsPhase *>(sm + RS_OFF_PH);
    uint64_t *w_full = reinterpret_cast<uint64_t *>(sm + RS_OFF_BAR), *w_empty = w_full + RS_NS, *a_full = w_empty + RS_NS, *acc_full = a_full + 1;
    uint64_t *s_full = acc_full + 1;                                          // [8 worker warps][3 slots]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_full + 24);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, G = gridDim.x;
    const int n_ph = plan->n_phases, B = plan->B;

    {   // phase table -> shared memory
        const uint4 *src = reinterpret_cast<const uint4 *>(plan->ph);
        uint4 *dst = reinterpret_cast<uint4 *>(phs);
        const int n16 = n_ph * (int)(sizeof(RsPhase) / 16);
        for (int i = tid; i < n16; i += RS_THREADS) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        if (smem_u32(sm) & 1023u) { printf("cpmusic: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < RS_NS; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
        mbar_init(a_full, 8);
        mbar_init(acc_full, 1);
        for (int i = 0; i < 24; ++i) mbar_init(s_full + i, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<32>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================================================================ weight producer
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int st = 0; st < n_steps; ++st)
                for (int pi = 0; pi < n_ph; ++pi) {
                    const RsPhase p = phs[pi];                      // a private copy: stores through generic pointers cannot alias it
                    if (p.type != PH_GEMM) continue;
                    const int tiles = p.n_wtiles * p.n_rtiles, KB = (p.K + 63) >> 6, per = p.BW == 64 ? 2 : 1;      // k-blocks per 16 KB stage
                    for (int t = cta; t < tiles; t += G) {
                        const int n0 = (t % p.n_wtiles) * p.BW;
                        for (int kb = 0; kb < KB; kb += per) {
                            const int nk = min(per, KB - kb);
                            mbar_wait(w_empty + s, ph ^ 1);
                            mbar_expect_tx(w_full + s, (uint32_t)(nk * p.BW) * 128u);
                            for (int j = 0; j < nk; ++j) tma_load_2d(sm + s * RS_STAGE + j * 8192, &plan->tm[pi], w_full + s, (kb + j) * 64, n0);
                            if (++s == RS_NS) { s = 0; ph ^= 1; }
                        }
                    }
                }
        }
    } else if (warp == 1) {
        // ================================================================ UMMA issuer
        if (lane == 0) {
            uint32_t s = 0, ph = 0, tile_par = 0;
            const uint64_t dW0 = smem_desc_sw128(smem_u32(sm)), dA0 = smem_desc_sw128(smem_u32(sa));
            for (int st = 0; st < n_steps; ++st)
                for (int pi = 0; pi < n_ph; ++pi) {
                    const RsPhase p = phs[pi];                      // a private copy: stores through generic pointers cannot alias it
                    if (p.type != PH_GEMM) continue;
                    const int tiles = p.n_wtiles * p.n_rtiles, KB = (p.K + 63) >> 6;
                    const uint32_t idesc = idesc_bf16(p.BW, p.BR, false, false);
                    for (int t = cta; t < tiles; t += G) {
                        mbar_wait(a_full, tile_par);              // activation tile staged (and the previous accumulator drained)
                        tc_fence_after();
                        if (timing && st < RS_TIMED_STEPS) timing[((int64_t)cta * RS_TIMED_STEPS * n_ph + st * n_ph + pi) * 8 + 6] = globaltimer_ns();
                        const int per = p.BW == 64 ? 2 : 1;
                        const uint64_t a_step = (uint64_t)((p.BR * 128) >> 4);
                        uint64_t dA = dA0;
                        for (int kb = 0; kb < KB; kb += per) {
                            mbar_wait(w_full + s, ph);
                            tc_fence_after();
                            const uint64_t dW = dW0 + (uint64_t)(s * (RS_STAGE >> 4));
                            mma_ss_kblock(tmem, dW, dA, idesc, kb > 0 ? 1u : 0u);
                            dA += a_step;
                            if (per == 2 && kb + 1 < KB) {
                                mma_ss_kblock(tmem, dW + (8192 >> 4), dA, idesc, 1u);
                                dA += a_step;
                            }
                            mma_commit(w_empty + s);
                            if (++s == RS_NS) { s = 0; ph ^= 1; }
                        }
                        mma_commit(acc_full);
                        if (timing && st < RS_TIMED_STEPS) timing[((int64_t)cta * RS_TIMED_STEPS * n_ph + st * n_ph + pi) * 8 + 7] = globaltimer_ns();
                        tile_par ^= 1;
                    }
                }
        }
    } else {
        // ================================================================ workers
        const int wtid = tid - 64, ww = wtid >> 5;
        const int quarter = warp & 3, half = (warp - 2) >> 2;     // TMEM lane quarter this warp may read; which half of the song columns
        unsigned long long target = 0;
        if (wtid == 0) target = (ld_acquire_u64(plan->barrier) / (unsigned)G) * (unsigned)G;
        const int step0 = __ldcg(plan->step_dev);
        uint32_t tile_par = 0, scnt = 0;
        for (int st = 0; st < n_steps; ++st) {
            const int step = step0 + st;
            int pos = plan->true_positions ? step : 0;
            pos = pos < plan->pe_len ? pos : plan->pe_len - 1;
            for (int pi = 0; pi < n_ph; ++pi) {
                const RsPhase p = phs[pi];                      // a private copy: stores through generic pointers cannot alias it
                if (p.type == PH_GEMM) {
                    const int tiles = p.n_wtiles * p.n_rtiles;
                    for (int t = cta; t < tiles; t += G) {
                        const int wt = t % p.n_wtiles, row0 = (t / p.n_wtiles) * p.BR;
                        if (p.pro == PRO_PLAIN) stage_plain(sa, p, row0, B, wtid);
                        else if (p.pro == PRO_EMBED) stage_embed(sa, p, plan, row0, B, wtid);
                        else stage_ln(sa, p, row0, B, ww, lane, plan->ln_eps, wt == 0 && p.xout != nullptr,
                                      (timing && wtid == 0 && st < RS_TIMED_STEPS) ? timing + ((int64_t)cta * RS_TIMED_STEPS * n_ph + st * n_ph + pi) * 8 : nullptr);
                        fence_proxy_async();                      // generic-proxy writes -> visible to the UMMA's async-proxy reads
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(a_full);
                        if (timing && wtid == 0 && st < RS_TIMED_STEPS) timing[((int64_t)cta * RS_TIMED_STEPS * n_ph + st * n_ph + pi) * 8 + 2] = globaltimer_ns();
                        // ---- epilogue (its constants are fetched while the UMMAs run)
                        const int n = wt * p.BW + (p.BW == 64 ? 16 * quarter + lane : 32 * quarter + lane);
                        const bool n_ok = (p.BW == 128 || lane < 16) && n < p.N;
                        const float bias = (n_ok && p.bias) ? __ldg(p.bias + n) : 0.f;
                        const float pe = (n_ok && p.epi == EPI_PE) ? __ldg(plan->pe + (int64_t)pos * p.N + n) : 0.f;
                        mbar_wait(acc_full, tile_par);
                        if (timing && wtid == 0 && st < RS_TIMED_STEPS) timing[((int64_t)cta * RS_TIMED_STEPS * n_ph + st * n_ph + pi) * 8 + 3] = globaltimer_ns();
                        tile_par ^= 1;
                        tc_fence_after();
                        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16);
                        if (p.BR == 32) {
                            uint32_t r[16];
                            tmem_ld16(taddr + half * 16, r);
                            tmem_ld_wait();
                            epilogue_cols<16>(r, p, n, n_ok, row0, half * 16, B, bias, pe);
                        } else {
                            uint32_t r[8];
                            tmem_ld8(taddr + half * 8, r);
                            tmem_ld_wait();
                            epilogue_cols<8>(r, p, n, n_ok, row0, half * 8, B, bias, pe);
                        }
                        tc_fence_before();
                    }
                } else if (p.type == PH_STATE) {
                    state_phase(sa, s_full, scnt, p, plan, cta, G, ww, lane);
                } else {
                    // ---- sample: one warp per (song, attribute)
                    const SegParams &sp = plan->seg;
                    float *buf = reinterpret_cast<float *>(sa) + ww * 512, *pr = buf + 256;
                    const int items = B * sp.n_attr;
                    for (int it = cta * 8 + ww; it < items; it += G * 8) {
                        const int r = it / sp.n_attr, a = it - r * sp.n_attr, w = sp.seg[a + 1] - sp.seg[a];
                        const __nv_bfloat16 *row = plan->logits + (int64_t)r * plan->logits_ld + sp.seg[a];
                        __syncwarp();
                        for (int i = lane; i < w; i += 32) buf[i] = ldcg_bf16(row + i);
                        ArgMax am;
                        float lse;
                        segment_stats(buf, w, lane, am, lse);
                        __syncwarp();
                        int tok = am.i;
                        if (!plan->greedy)
                            tok = sample_segment<8>(buf, pr, w, lane, am, sp.temperature[a], sp.top_p[a], plan->seed, (uint64_t)(plan->seq_base + r), step, a);
                        if (lane == 0) {
                            const float lp = buf[tok] - lse;
                            plan->cur[it] = tok;
                            plan->logp[it] = lp;
                            if (step < plan->max_steps) {
                                if (plan->hist_tok) plan->hist_tok[(int64_t)step * items + it] = tok;
                                if (plan->hist_logp) plan->hist_logp[(int64_t)step * items + it] = lp;
                            }
                        }
                    }
                    if (cta == 0 && wtid == 0) *plan->step_dev = step + 1;
                }
                if (timing && wtid == 0 && st < RS_TIMED_STEPS) {
                    timing[((int64_t)cta * RS_TIMED_STEPS * n_ph + st * n_ph + pi) * 8] = globaltimer_ns();
                    if (p.type != PH_GEMM) timing[((int64_t)cta * RS_TIMED_STEPS * n_ph + st * n_ph + pi) * 8 + 2] = (unsigned long long)clock64();   // SM clock estimate
                }
                if (!(st == n_steps - 1 && pi == n_ph - 1)) grid_barrier(plan->barrier, target, wtid);
                if (timing && wtid == 0 && st < RS_TIMED_STEPS) timing[((int64_t)cta * RS_TIMED_STEPS * n_ph + st * n_ph + pi) * 8 + 1] = globaltimer_ns();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<32>(tmem);
}

// ---------------------------------------------------------------- host
unsigned long long *g_rs_timing = nullptr;

struct Handle {
    RsPlan *plan_dev;
    int n_phases, grid;
};

// tile shape of one Linear stage: BR songs x BW weight rows; as many tiles as fit one wave of `G` CTAs, ties -> least operand bytes
void pick_tiles(int B, int N, int K, int G, RsPhase &p) {
    int best_tiles = -1, best_bytes = 0;
    for (int BR : {16, 32})
        for (int BW : {64, 128}) {
            if ((int64_t)BR * (((K + 63) / 64) * 64) * 2 > RS_A_BYTES) continue;
            const int rt = (B + BR - 1) / BR, wt = (N + BW - 1) / BW, tiles = rt * wt, bytes = (BR + BW) * K;
            const bool fits = tiles <= G, best_fits = best_tiles >= 0 && best_tiles <= G;
            bool better;
            if (best_tiles < 0) better = true;
            else if (fits != best_fits) better = fits;
            else if (fits) better = tiles > best_tiles || (tiles == best_tiles && bytes < best_bytes);
            else better = bytes * tiles < best_bytes * best_tiles;
            if (better) { best_tiles = tiles; best_bytes = bytes; p.BR = BR; p.BW = BW; p.n_rtiles = rt; p.n_wtiles = wt; }
        }
}

}  // namespace
}  // namespace cpm

using namespace cpm;

extern "C" int64_t cpm_rollout_plan_bytes(void) { return (int64_t)sizeof(RsPlan); }

extern "C" int cpm_rollout_create(const CpmRolloutConfig *c, void *plan_dev, void **handle_out) {
    CPM_REQUIRE(c && plan_dev && handle_out, CPM_ERR_NULL, "rollout_create: NULL argument");
    CPM_REQUIRE((reinterpret_cast<uintptr_t>(plan_dev) & 255u) == 0, CPM_ERR_BAD_ALIGN, "rollout_create: plan_dev must be 256-byte aligned");
    const int d = c->d_model, B = c->batch;
    CPM_REQUIRE(B > 0 && c->n_layers > 0 && c->n_layers <= CPM_ROLLOUT_MAX_LAYERS && c->n_attr >= 1 && c->n_attr <= CPM_MAX_ATTR, CPM_ERR_BAD_SHAPE,
                "rollout_create: batch=%d n_layers=%d n_attr=%d", B, c->n_layers, c->n_attr);
    if (d <= 0 || d > 512 || d % 64 || c->n_heads * 64 != d || c->d_ff <= 0 || c->d_ff % 64 || c->d_ff > 2048)
        return fail(CPM_ERR_UNSUPPORTED, "rollout_create: d_model=%d n_heads=%d d_ff=%d (needs 64-wide heads, d_model <= 512, d_ff <= 2048, multiples of 64)",
                    d, c->n_heads, c->d_ff);
    if (2 + 5 * c->n_layers + 2 > RS_MAX_PH) return fail(CPM_ERR_UNSUPPORTED, "rollout_create: %d layers need more than %d stages", c->n_layers, RS_MAX_PH);
    int k_in = 0;
    for (int a = 0; a < c->n_attr; ++a) {
        CPM_REQUIRE(c->tables[a] && c->emb[a] > 0 && c->emb[a] % 8 == 0 && c->n_tokens[a] > 0, CPM_ERR_BAD_SHAPE, "rollout_create: attribute %d", a);
        if (c->seg[a + 1] - c->seg[a] > 256 || c->seg[a + 1] - c->seg[a] != c->n_tokens[a])
            return fail(CPM_ERR_UNSUPPORTED, "rollout_create: attribute %d has %d classes (<= 256 and equal to its vocabulary)", a, c->seg[a + 1] - c->seg[a]);
        k_in += c->emb[a];
    }
    if ((int64_t)16 * ((k_in + 63) / 64 * 64) * 2 > RS_A_BYTES) return fail(CPM_ERR_UNSUPPORTED, "rollout_create: embedding width %d", k_in);
    CPM_REQUIRE(c->logits_ld >= c->seg[c->n_attr] && c->logits_ld % 8 == 0, CPM_ERR_BAD_SHAPE, "rollout_create: logits_ld=%d", c->logits_ld);
    CPM_REQUIRE(c->w_in && c->b_in && c->pe && c->pe_len > 0 && c->lnf_g && c->lnf_b && c->w_heads && c->b_heads && c->cur && c->logp && c->step_dev &&
                    c->x0 && c->x1 && c->y && c->qkv && c->attn && c->g && c->logits && c->barrier,
                CPM_ERR_NULL, "rollout_create: a required pointer is NULL");
    int dev = 0, sms = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || sms <= 0) return fail(CPM_ERR_UNSUPPORTED, "rollout_create: the device does not support cooperative launches");
    cudaError_t e = cudaFuncSetAttribute(rollout_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM);
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "rollout_create: shared-memory attribute: %s", cudaGetErrorString(e));
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_step_kernel, RS_THREADS, RS_SMEM);
    if (e != cudaSuccess || per_sm < 1) return fail(CPM_ERR_CUDA, "rollout_create: the step kernel does not fit an SM (%s)", cudaGetErrorString(e));
    const int G = sms;

    RsPlan *pl = new (std::nothrow) RsPlan();
    CPM_REQUIRE(pl, CPM_ERR_CUDA, "rollout_create: out of host memory");
    memset(pl, 0, sizeof(RsPlan));
    int np = 0, rc = CPM_OK;
    auto bf = [](const void *p) { return (const __nv_bfloat16 *)p; };
    auto gemm = [&](const void *W, const float *bias, int N, int K, int tm_rows, int pro, int epi, const void *A, int lda, void *D, int ldd) -> RsPhase & {
        RsPhase &p = pl->ph[np];
        p.type = PH_GEMM; p.N = N; p.K = K; p.pro = pro; p.epi = epi; p.A = bf(A); p.lda = lda; p.bias = bias; p.D = (__nv_bfloat16 *)D; p.ldd = ldd;
        pick_tiles(B, N, K, G, p);
        if (!W || !aligned16(W)) rc = fail(CPM_ERR_BAD_ALIGN, "rollout_create: a weight matrix is NULL or not 16-byte aligned");
        else if (rc == CPM_OK) rc = make_tmap_bf16_2d(&pl->tm[np], W, (uint64_t)K, (uint64_t)tm_rows, (uint64_t)K, (uint32_t)p.BW);
        ++np;
        return p;
    };
    __nv_bfloat16 *x0 = (__nv_bfloat16 *)c->x0, *x1 = (__nv_bfloat16 *)c->x1, *y = (__nv_bfloat16 *)c->y;
    gemm(c->w_in, c->b_in, d, k_in, d, PRO_EMBED, EPI_PE, nullptr, 0, x0, d);
    for (int l = 0; l < c->n_layers; ++l) {
        const CpmRolloutLayer &L = c->layer[l];
        if (!(L.b_qkv && L.b_out && L.ln1_g && L.ln1_b && L.b_ff1 && L.b_ff2 && L.ln2_g && L.ln2_b && L.S && L.Z)) {
            rc = fail(CPM_ERR_NULL, "rollout_create: layer %d has a NULL pointer", l);
            break;
        }
        if (l == 0) gemm(L.w_qkv, L.b_qkv, 3 * d, d, 3 * d, PRO_PLAIN, EPI_STORE, x0, d, c->qkv, 3 * d);
        else {
            RsPhase &p = gemm(L.w_qkv, L.b_qkv, 3 * d, d, 3 * d, PRO_LN, EPI_STORE, y, d, c->qkv, 3 * d);      // norm2 of layer l - 1
            p.g1 = c->layer[l - 1].ln2_g; p.b1 = c->layer[l - 1].ln2_b; p.R = x1; p.ldr = d; p.xout = x0;
        }
        {
            RsPhase &p = pl->ph[np++];
            p.type = PH_STATE; p.S = L.S; p.Z = L.Z;
        }
        gemm(L.w_out, L.b_out, d, d, d, PRO_PLAIN, EPI_STORE, c->attn, d, y, d);
        {
            RsPhase &p = gemm(L.w_ff1, L.b_ff1, c->d_ff, d, c->d_ff, PRO_LN, EPI_GELU, y, d, c->g, c->d_ff);                 // norm1(x0 + out-projection)
            p.g1 = L.ln1_g; p.b1 = L.ln1_b; p.R = x0; p.ldr = d; p.xout = x1;
        }
        gemm(L.w_ff2, L.b_ff2, d, c->d_ff, d, PRO_PLAIN, EPI_STORE, c->g, c->d_ff, y, d);
    }
    if (rc == CPM_OK) {
        const CpmRolloutLayer &L = c->layer[c->n_layers - 1];
        RsPhase &p = gemm(c->w_heads, c->b_heads, c->seg[c->n_attr], d, c->logits_ld, PRO_LN2, EPI_STORE, y, d, c->logits, c->logits_ld);
        p.g1 = L.ln2_g; p.b1 = L.ln2_b; p.g2 = c->lnf_g; p.b2 = c->lnf_b; p.R = x1; p.ldr = d;
        pl->ph[np++].type = PH_SAMPLE;
    }
    if (rc != CPM_OK) { delete pl; return rc; }
    pl->n_phases = np; pl->B = B; pl->H = c->n_heads; pl->d_model = d; pl->logits_ld = c->logits_ld; pl->pe_len = c->pe_len;
    pl->true_positions = c->true_positions; pl->greedy = c->greedy; pl->max_steps = c->max_steps;
    pl->ln_eps = c->ln_eps; pl->attn_eps = c->attn_eps; pl->seed = c->seed; pl->seq_base = c->seq_base;
    pl->emb.n_attr = c->n_attr;
    for (int a = 0; a < c->n_attr; ++a) {
        pl->emb.tables[a] = c->tables[a]; pl->emb.n_tokens[a] = c->n_tokens[a]; pl->emb.emb[a] = c->emb[a];
        pl->emb.off[a + 1] = pl->emb.off[a] + c->emb[a];
        pl->emb.scale[a] = sqrtf((float)c->emb[a]);
        pl->seg.seg[a] = c->seg[a]; pl->seg.temperature[a] = c->temperature[a]; pl->seg.top_p[a] = c->top_p[a];
    }
    pl->seg.seg[c->n_attr] = c->seg[c->n_attr];
    pl->seg.n_attr = c->n_attr;
    pl->pe = c->pe; pl->cur = c->cur; pl->logp = c->logp; pl->hist_tok = c->hist_tok; pl->hist_logp = c->hist_logp; pl->step_dev = c->step_dev;
    pl->qkv = bf(c->qkv); pl->attn = (__nv_bfloat16 *)c->attn; pl->logits = bf(c->logits); pl->barrier = c->barrier; pl->err_flag = c->err_flag;
    e = cudaMemcpy(plan_dev, pl, sizeof(RsPlan), cudaMemcpyHostToDevice);
    delete pl;
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "rollout_create: plan upload: %s", cudaGetErrorString(e));
    Handle *h = new (std::nothrow) Handle();
    CPM_REQUIRE(h, CPM_ERR_CUDA, "rollout_create: out of host memory");
    h->plan_dev = (RsPlan *)plan_dev; h->n_phases = np; h->grid = G;
    *handle_out = h;
    return CPM_OK;
}

extern "C" int cpm_rollout_run(void *handle, int n_steps, void *stream) {
    CPM_REQUIRE(handle, CPM_ERR_NULL, "rollout_run: NULL handle");
    CPM_REQUIRE(n_steps > 0, CPM_ERR_BAD_SHAPE, "rollout_run: n_steps=%d", n_steps);
    Handle *h = (Handle *)handle;
    const RsPlan *plan = h->plan_dev;
    unsigned long long *timing = g_rs_timing;
    void *args[] = {(void *)&plan, (void *)&n_steps, (void *)&timing};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)rollout_step_kernel, dim3(h->grid), dim3(RS_THREADS), args, RS_SMEM, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(CPM_ERR_CUDA, "rollout_run launch: %s", cudaGetErrorString(e));
    return CPM_OK;
}

extern "C" int cpm_debug_rollout_timing(void *device_buffer) {
    g_rs_timing = (unsigned long long *)device_buffer;
    return CPM_OK;
}

extern "C" int cpm_rollout_phases(void *handle) { return handle ? ((Handle *)handle)->n_phases : 0; }

extern "C" int cpm_rollout_destroy(void *handle) {
    delete (Handle *)handle;
    return CPM_OK;
}
