// rollout.cu -- qs_rollout: T fused control ticks per launch.
#include "qs_common.cuh"
#include <cuda_fp16.h>

using namespace qsi;

namespace {

// ---------------------------------------------------------------------------------------------------------
// Multi-tick rollout: the fused control tick in a loop.  Drone state lives in registers, the CTA's observation rows in a
// shared-memory window that slides by one action per tick (new_flat[j] = old_flat[j + A], so "shifting the history" is
// `base += A`); per tick the kernel reads the action and writes the rows, reward and flags.  Bit-identical to T calls of
// qs_step: the state is rounded to its float32 plane representation at every tick boundary exactly like store/load.
// ---------------------------------------------------------------------------------------------------------
struct RolloutArgs {
    QsParams P;
    QsState st;
    QsRolloutIO io;
    QsPolicy pol;        // copy of *io.policy (device pointers inside) when POLICY
    int act_type, task, n_envs, D, substeps, N, A, obs_dim, tpb;
    unsigned effects, flags;
    int stage_mode, cap;
};

// ---- on-device policy: SB3-MlpPolicy-shaped MLP on the tensor cores, fp32-accurate ------------------------------------------------
// One warp takes 16 aviaries (one m-tile) through the whole network: Y[16][64] = X[16][K] W[K][64] per layer as mma.sync.m16n8k16
// F16 tiles with a two-term split of both operands: x = x_hi + x_lo, w = w_hi + w_lo with x_hi = fp16(x) and
// x_lo' = fp16(2^11 (x - x_hi)) (the scaling keeps the remainder out of the fp16 subnormals), and
//     x w ~ x_hi w_hi + 2^-11 (x_hi w_lo' + x_lo' w_hi)
// with fp32 accumulation in two accumulator sets -- relative error ~2^-21 per product, i.e. fp32-level (a plain F16/TF32/BF16 mma has
// 2^-11 / 2^-8 and fails the 1e-5 parity with the fp32 torch network).  FP16 and TF32 carry the same 11 significant bits, but one
// m16n8k16 F16 instruction does twice the work of an m16n8k8 TF32 one at the same issue rate (8 cycles per SM sub-partition,
// tools/mma_rate.cu): 3 instead of 6 mma per 16x8x16 block.
//  * Weights: split once on the host (MlpPolicy) and stored in FRAGMENT ORDER, [k-step][n-tile][lane] x 16 bytes = the lane's
//    {b0 hi, b1 hi, b0 lo', b1 lo'} registers, so a B fragment is one fully coalesced 128-bit load (L1-resident: the CTAs use
//    ~160 KB of the SM's 256 KB as shared memory, the actor's 55 KB of weights stay in the rest).
//  * First layer: the observation rows are read from the shared-memory window as 128-bit loads -- lane t supplies elements
//    4t .. 4t+3 of each 16-wide k-step instead of the canonical {2t, 2t+1, 2t+8, 2t+9}; W1 is packed with the same permutation of
//    k, so the product is unchanged -- and split per fragment (conversions saturate to +-65504: an observation beyond that range
//    acts like a clipped one).
//  * Hidden layers never leave the registers: the accumulator fragment of n-tiles (2j, 2j+1) IS the A fragment of k-step j of the
//    next layer (tanh, split, pack) -- no shared-memory round trip, no barrier between layers.
// K = 144 / 64 and M = 32 rows per CTA are far below a tcgen05 tile (M = 128 per CTA, operands through shared-memory descriptors,
// accumulators in TMEM) -- see DESIGN.md 4.1b for the numbers.
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

__device__ __forceinline__ unsigned pack_f16x2_sat(float lo_elem, float hi_elem) {
    unsigned r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
// (x0, x1) -> packed halves of the leading parts and of the scaled remainders
__device__ __forceinline__ void f16_split2(float x0, float x1, unsigned& hi, unsigned& lo) {
    hi = pack_f16x2_sat(x0, x1);
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    lo = pack_f16x2_sat((x0 - f.x) * kLoScale, (x1 - f.y) * kLoScale);
}
__device__ __forceinline__ void mma_f16(float c[4], const unsigned a[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// tanh(x) = 1 - 2 / (exp(2x) + 1) on the special-function unit: absolute error ~2e-7 (the libm tanhf costs ~4x the instructions)
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

// G n-tiles from n0 against one A fragment: three sweeps, so that no mma waits for the one issued just before it
template <int G>
__device__ __forceinline__ void mma_group(float (*c)[4], float (*cx)[4], const unsigned (&ah)[4], const unsigned (&al)[4], const uint4 (&b)[G]) {
#pragma unroll
    for (int n = 0; n < G; ++n) mma_f16(cx[n], al, b[n].x, b[n].y);          // x_lo' w_hi
#pragma unroll
    for (int n = 0; n < G; ++n) mma_f16(c[n], ah, b[n].x, b[n].y);           // x_hi  w_hi
#pragma unroll
    for (int n = 0; n < G; ++n) mma_f16(cx[n], ah, b[n].z, b[n].w);          // x_hi  w_lo'
}
template <int G>
__device__ __forceinline__ void load_bfrag(uint4 (&b)[G], const uint4* __restrict__ W) {
#pragma unroll
    for (int n = 0; n < G; ++n) b[n] = __ldg(W + 32 * n);
}
template <int NT>
__device__ __forceinline__ void init_acc(float (&c)[NT][4], float (&cx)[NT][4], const float* __restrict__ bias, int t) {
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const float2 b = __ldg(reinterpret_cast<const float2*>(bias + 8 * n + 2 * t));
        c[n][0] = b.x; c[n][1] = b.y; c[n][2] = b.x; c[n][3] = b.y;
        cx[n][0] = 0.f; cx[n][1] = 0.f; cx[n][2] = 0.f; cx[n][3] = 0.f;
    }
}
// accumulators of a 64-unit hidden layer -> tanh -> the 4 k-steps of A fragments of the next layer
__device__ __forceinline__ void hidden_to_afrag(const float (&c)[8][4], const float (&cx)[8][4], unsigned (&hh)[4][4], unsigned (&hl)[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = 2 * j + h;
            const float v0 = tanh_fast(fmaf(cx[n][0], kLoInv, c[n][0])), v1 = tanh_fast(fmaf(cx[n][1], kLoInv, c[n][1]));
            const float v2 = tanh_fast(fmaf(cx[n][2], kLoInv, c[n][2])), v3 = tanh_fast(fmaf(cx[n][3], kLoInv, c[n][3]));
            f16_split2(v0, v1, hh[j][2 * h], hl[j][2 * h]);                  // row g
            f16_split2(v2, v3, hh[j][2 * h + 1], hl[j][2 * h + 1]);          // row g + 8
        }
}

// One network (in -> 64 tanh -> 64 tanh -> 8 nt3 outputs) for 16 aviaries, by one warp.  x_s: rows of in_dim fp32 values in shared
// memory (rows past n_rows repeat the last row; their outputs are not stored); W1/W2/W3: fragment-ordered weights (see above; W1 with
// the 4t permutation); out_s rows have stride 8 nt3 floats (padded outputs).
__device__ __forceinline__ void warp_net(const float* x_s, int in_dim, const uint4* __restrict__ W1, const uint4* __restrict__ W2,
                                         const uint4* __restrict__ W3, const float* b1, const float* b2, const float* b3, int nt3, float* out_s,
                                         int n_rows, int lane) {
    const int g = lane >> 2, t = lane & 3;
    unsigned hh[4][4], hl[4][4];
    {   // ---- layer 1: K = in_dim from shared memory ----
        float c[8][4], cx[8][4];
        init_acc<8>(c, cx, b1, t);
        const int last = n_rows - 1;
        const float* p0 = x_s + (size_t)(g < last ? g : last) * in_dim + 4 * t;
        const float* p1 = x_s + (size_t)(g + 8 < last ? g + 8 : last) * in_dim + 4 * t;
        const bool vec4 = ((reinterpret_cast<size_t>(x_s) & 15) == 0) && ((in_dim & 3) == 0);
        // four rotating B buffers of 2 n-tiles: each is refilled with the next k-step's tiles right after its mma group, i.e. three
        // groups (18 mma) plus the next fragment split ahead of its use
        const uint4* w = W1 + lane;
        uint4 b0[2], b1[2], b2[2], b3[2];
        load_bfrag<2>(b0, w); load_bfrag<2>(b1, w + 2 * 32); load_bfrag<2>(b2, w + 4 * 32); load_bfrag<2>(b3, w + 6 * 32);
#pragma unroll 1
        for (int k0 = 0; k0 < in_dim; k0 += 16, w += 8 * 32) {
            unsigned ah[4], al[4];
            if (vec4 && k0 + 16 <= in_dim) {
                const float4 u = *reinterpret_cast<const float4*>(p0 + k0), v = *reinterpret_cast<const float4*>(p1 + k0);
                f16_split2(u.x, u.y, ah[0], al[0]); f16_split2(v.x, v.y, ah[1], al[1]);
                f16_split2(u.z, u.w, ah[2], al[2]); f16_split2(v.z, v.w, ah[3], al[3]);
            } else {                                         // unaligned rows (odd action width) or the ragged last k-step: elements past
                const int ka = k0 + 4 * t;                   // in_dim are zeros (their weights are zero rows)
                const float* q0 = p0 + k0; const float* q1 = p1 + k0;
                f16_split2(ka < in_dim ? q0[0] : 0.f, ka + 1 < in_dim ? q0[1] : 0.f, ah[0], al[0]);
                f16_split2(ka < in_dim ? q1[0] : 0.f, ka + 1 < in_dim ? q1[1] : 0.f, ah[1], al[1]);
                f16_split2(ka + 2 < in_dim ? q0[2] : 0.f, ka + 3 < in_dim ? q0[3] : 0.f, ah[2], al[2]);
                f16_split2(ka + 2 < in_dim ? q1[2] : 0.f, ka + 3 < in_dim ? q1[3] : 0.f, ah[3], al[3]);
            }
            const bool more = k0 + 16 < in_dim;
            mma_group<2>(c, cx, ah, al, b0);         if (more) load_bfrag<2>(b0, w + 8 * 32);
            mma_group<2>(c + 2, cx + 2, ah, al, b1); if (more) load_bfrag<2>(b1, w + 10 * 32);
            mma_group<2>(c + 4, cx + 4, ah, al, b2); if (more) load_bfrag<2>(b2, w + 12 * 32);
            mma_group<2>(c + 6, cx + 6, ah, al, b3); if (more) load_bfrag<2>(b3, w + 14 * 32);
        }
        hidden_to_afrag(c, cx, hh, hl);
    }
    {   // ---- layer 2: K = 64 from registers ----
        float c[8][4], cx[8][4];
        init_acc<8>(c, cx, b2, t);
        const uint4* w = W2 + lane;
        uint4 ba[2], bb[2];
        load_bfrag<2>(ba, w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int n = 0; n < 8; n += 4) {
                load_bfrag<2>(bb, w + (8 * j + n + 2) * 32);
                mma_group<2>(c + n, cx + n, hh[j], hl[j], ba);
                if (8 * j + n + 4 < 32) load_bfrag<2>(ba, w + (8 * j + n + 4) * 32);
                mma_group<2>(c + n + 2, cx + n + 2, hh[j], hl[j], bb);
            }
        }
        hidden_to_afrag(c, cx, hh, hl);
    }
    // ---- layer 3: K = 64 from registers, one n-tile of 8 outputs at a time ----
    const int ost = 8 * nt3;
    for (int n = 0; n < nt3; ++n) {
        float c[1][4], cx[1][4];
        init_acc<1>(c, cx, b3 + 8 * n, t);
        uint4 b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = __ldg(W3 + ((size_t)j * nt3 + n) * 32 + lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mma_f16(cx[0], hl[j], b[j].x, b[j].y); mma_f16(c[0], hh[j], b[j].x, b[j].y); mma_f16(cx[0], hh[j], b[j].z, b[j].w);
        }
        float* o0 = out_s + (size_t)g * ost + 8 * n + 2 * t;
        if (g < n_rows) *reinterpret_cast<float2*>(o0) = make_float2(fmaf(cx[0][0], kLoInv, c[0][0]), fmaf(cx[0][1], kLoInv, c[0][1]));
        if (g + 8 < n_rows) *reinterpret_cast<float2*>(o0 + 8 * ost) = make_float2(fmaf(cx[0][2], kLoInv, c[0][2]), fmaf(cx[0][3], kLoInv, c[0][3]));
    }
}

// The policy part of one tick for the whole CTA: actor (and critic) over the CTA's aviaries in tiles of 16; work item i = net x tile
// goes to warp i & 1.  Deliberately NOT inlined: inside the tick loop its ~120 live registers made the compiler spill the drone state
// in the middle of the physics substeps; as a call, the state is saved once per tick around it.  `parked` is not touched: the caller
// hands over the address of its drone state so that the state demonstrably lives in local memory across the call (one store + load
// per tick) instead of being spilled piecemeal inside the substep loop.
__device__ __noinline__ void policy_forward(const RolloutArgs& a, const float* base, float* mean_s, float* val_s, int n_av, int t, qs::Drone* parked) {
    if (n_av < 0) parked->px = 0.0;                          // never taken; keeps the hand-over opaque to the optimiser
    const int warp = t >> 5, lane = t & 31, ost = 8 * a.pol.nt3;
    const int tiles = (n_av + 15) >> 4, items = a.pol.vw1 ? 2 * tiles : tiles;
    for (int i = warp; i < items; i += 2) {
        const bool critic = i >= tiles;
        const int r0 = 16 * (critic ? i - tiles : i);
        const float* x0 = base + (size_t)r0 * a.pol.in_dim;
        if (!critic)
            warp_net(x0, a.pol.in_dim, reinterpret_cast<const uint4*>(a.pol.w1), reinterpret_cast<const uint4*>(a.pol.w2), reinterpret_cast<const uint4*>(a.pol.w3),
                     a.pol.b1, a.pol.b2, a.pol.b3, a.pol.nt3, mean_s + (size_t)r0 * ost, n_av - r0, lane);
        else
            warp_net(x0, a.pol.in_dim, reinterpret_cast<const uint4*>(a.pol.vw1), reinterpret_cast<const uint4*>(a.pol.vw2), reinterpret_cast<const uint4*>(a.pol.vw3),
                     a.pol.vb1, a.pol.vb2, a.pol.vb3, 1, val_s + (size_t)r0 * 8, n_av - r0, lane);
    }
    __syncthreads();                                         // means / values of every aviary of the CTA are in shared memory
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u32_to_pm1(unsigned u) { return (float)(u >> 8) * (1.0f / 8388608.0f) - 1.0f; }   // [-1, 1)

// what store_drone + load_drone do to the state between two ticks: the quaternion is renormalised, nothing is rounded
// (the planes are float64), so T fused ticks equal T calls of qs_step bit for bit
__device__ __forceinline__ void round_to_planes(qs::Drone& d) {
    const double inv = rsqrt(qs::quat_norm2(d.qx, d.qy, d.qz, d.qw));
    d.qx = __dmul_rn(d.qx, inv); d.qy = __dmul_rn(d.qy, inv); d.qz = __dmul_rn(d.qz, inv); d.qw = __dmul_rn(d.qw, inv);
}

template <int EFF, bool PIDACT, bool POLICY>
__global__ void __launch_bounds__(POLICY ? 64 : kMaxTPB, POLICY ? 7 : 4) rollout_kernel(const __grid_constant__ RolloutArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const QsParams& P = a.P;
    const int tpb = a.tpb, D = a.D, A = a.A, od = a.obs_dim, T = a.io.T;
    const int t = threadIdx.x;
    const long long N = a.N, E = a.n_envs;
    const long long c0 = (long long)blockIdx.x * tpb;
    const long long i = c0 + t;
    const bool live = (t < tpb) && (i < N);
    const int rows = (int)((N - c0) < tpb ? (N - c0) : tpb);
    double* red_s = reinterpret_cast<double*>(smem_raw);                               // [tpb][2]
    const int cap = a.cap;
    // POLICY (no DYN+ effects, CTA of 64 drones): a compact fixed part -- red_s [64][2] doubles, oob [64], done [64], mbarrier --
    // so that 7 CTAs (one wave of 1024 CTAs on 148 SMs) fit into shared memory next to the window and the MLP scratch
    const size_t fixed = POLICY ? (size_t)(64 * 2 * 8 + 64 + 64 + 16 + 112) : smem_fixed(cap);       // 1280 for POLICY
    double* pos_s = red_s + (size_t)cap * 2;                                           // [tpb][3] (in-CTA downwash only)
    unsigned char* oob_s = POLICY ? reinterpret_cast<unsigned char*>(red_s + 128) : reinterpret_cast<unsigned char*>(pos_s + (size_t)cap * 3);
    unsigned char* done_s = oob_s + (POLICY ? 64 : cap);
    unsigned long long* bar_s = reinterpret_cast<unsigned long long*>(smem_raw + fixed - 16);
    float* stage_s = reinterpret_cast<float*>(smem_raw + fixed);                       // [tpb*od + (T+1)*A] sliding window
    // POLICY scratch after the window: padded action means [n_av][8 nt3]; padded values [n_av][8]; log-prob terms [64]
    float* pol_s = stage_s + ((((size_t)tpb * od + (size_t)(T + 1) * A) + 3) & ~(size_t)3);
    float* mean_s = pol_s;
    float* val_s = mean_s + (size_t)(64 / (D < 1 ? 1 : D)) * 8 * a.pol.nt3;
    float* lp_s = val_s + (size_t)(64 / (D < 1 ? 1 : D)) * 8;

    const long long e = live ? i / D : 0;
    const int le = t / D;
    const int dslot = (int)(i - e * D);
    const long long tbl = a.st.tables_per_env ? i : dslot;

    if (a.stage_mode == 1) {
        if (t == 0) mbar_init(bar_s, 1);
        __syncthreads();
    }
    qs::Drone d;
    qs::PidState pst = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double rpm_prev[4] = {0, 0, 0, 0};
    int sc = 0;
    if (live) {
        load_drone(a.st.planes, N, i, d);
        if ((EFF & QS_EFFECT_DRAG) && a.st.last_rpm) load_rpm(a.st.last_rpm, i, rpm_prev);
        if (PIDACT) load_pid(a.st.pid, N, i, pst);
        sc = a.st.step_counter[e];
    }
    if (a.stage_mode == 1) {
        if (t == 0) tma_bulk_g2s(stage_s, a.io.obs_init + c0 * od, (unsigned)(rows * od * 4), bar_s);
        mbar_wait(bar_s, 0);
    } else {
        const float* src = a.io.obs_init + c0 * od;
        for (int j = t; j < rows * od; j += blockDim.x) cp_async4(stage_s + j, src + j);
        cp_async_commit_wait_all();
    }
    __syncthreads();

    float* base = stage_s;                     // window start: rows of the observation BEFORE the current tick
    double rpm[4] = {0, 0, 0, 0};
    for (int k = 0; k < T; ++k) {
        // ---- this tick's action: caller-provided or generated on the device --------------------------------------
        float act[4] = {0.f, 0.f, 0.f, 0.f};
        float raw_act[4] = {0.f, 0.f, 0.f, 0.f};
        if (POLICY) {
            // the aviaries of this CTA: rows [le D, le D + D) of the window = one flattened observation of in_dim floats each
            const int n_av = rows / D;
            const int ost = 8 * a.pol.nt3;                        // padded width of the output rows in mean_s
            {
                qs::Drone parked = d;
                policy_forward(a, base, mean_s, val_s, n_av, t, &parked);
                d = parked;
            }
            float lp = 0.f;
            if (live) {
                const int od_out = a.pol.out_dim;
                for (int j = 0; j < A; ++j) {
                    const int idx = dslot * A + j;
                    const float ls = __ldg(a.pol.log_std + idx);
                    const float eps = a.pol.noise ? __ldg(a.pol.noise + ((long long)k * E + e) * od_out + idx) : 0.f;
                    const float r = fmaf(expf(ls), eps, mean_s[le * ost + idx]);
                    raw_act[j] = r;
                    act[j] = fminf(fmaxf(r, -1.f), 1.f);                             // the env clips to its action space
                    lp += -0.5f * eps * eps - ls - 0.91893853320467274f;            // log N(r; mean, std)
                }
                lp_s[t] = lp;
            }
            __syncthreads();
            if (live && dslot == 0) {
                float s_lp = 0.f;
                for (int q = 0; q < D; ++q) s_lp += lp_s[t + q];
                const long long oe = (long long)k * E + e;
                if (a.pol.logprob) a.pol.logprob[oe] = s_lp;
                if (a.pol.values && a.pol.vw1) a.pol.values[oe] = val_s[le * 8];
            }
        }
        if (live) {
            if (POLICY) {
                // (act / raw_act set above)
            } else if (a.io.actions) {
                const float* ap = a.io.actions + ((long long)k * N + i) * A;
                if (A == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(ap)); act[0] = v.x; act[1] = v.y; act[2] = v.z; act[3] = v.w; }
                else if (A == 3) { act[0] = __ldg(ap); act[1] = __ldg(ap + 1); act[2] = __ldg(ap + 2); }
                else act[0] = __ldg(ap);
            } else {
                const unsigned long long key = a.io.seed + 2ull * (unsigned long long)((a.io.tick0 + k) * N + i);
                const unsigned long long r0 = splitmix64(key), r1 = splitmix64(key + 1);
                act[0] = u32_to_pm1((unsigned)r0); act[1] = u32_to_pm1((unsigned)(r0 >> 32));
                act[2] = u32_to_pm1((unsigned)r1); act[3] = u32_to_pm1((unsigned)(r1 >> 32));
                if (A < 4) act[3] = 0.f;
                if (A < 3) { act[1] = 0.f; act[2] = 0.f; }
            }
            float* tail = base + (size_t)t * od + od;        // new action -> the A slots after my row (dead head of the next row)
            if (A == 4) *reinterpret_cast<float4*>(tail) = make_float4(act[0], act[1], act[2], act[3]);
            else if (A == 3) { tail[0] = act[0]; tail[1] = act[1]; tail[2] = act[2]; }
            else tail[0] = act[0];
            if (a.io.actions_out) {
                float* ao = a.io.actions_out + ((long long)k * N + i) * A;
                const float* av = POLICY ? raw_act : act;                        // PPO stores the unclipped sample
                if (A == 4) *reinterpret_cast<float4*>(ao) = make_float4(av[0], av[1], av[2], av[3]);
                else if (A == 3) { ao[0] = av[0]; ao[1] = av[1]; ao[2] = av[2]; }
                else ao[0] = av[0];
            }
        }
        // ---- physics ---------------------------------------------------------------------------------------------
        double R_last[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (live) {
            double cur_yaw = 0.0;
            if (a.act_type == QS_ACT_VEL) { double r_, p_; qs::quat_to_euler<false>(d.qx, d.qy, d.qz, d.qw, r_, p_, cur_yaw); }
            qs::decode_action<PIDACT>(P, a.act_type, act, d, cur_yaw, pst, rpm);
        }
        if (EFF & QS_EFFECT_DW) {
            for (int s = 0; s < a.substeps; ++s) {
                if (live) { pos_s[3 * t] = d.px; pos_s[3 * t + 1] = d.py; pos_s[3 * t + 2] = d.pz; }
                __syncthreads();
                if (live) {
                    double fz = 0.0;
                    const int b = le * D;
                    for (int q = 0; q < D; ++q) {
                        const double dz = pos_s[3 * (b + q) + 2] - d.pz;
                        const double dx = pos_s[3 * (b + q)] - d.px, dy = pos_s[3 * (b + q) + 1] - d.py;
                        const double dxy2 = dx * dx + dy * dy;
                        if (dz > 0.0 && dxy2 < 100.0) fz += qs::downwash_pair(P, dz, dxy2);
                    }
                    qs::dyn_tick<EFF>(P, d, rpm, s == 0 ? rpm_prev : rpm, fz, 1, R_last);
                }
                __syncthreads();
            }
        } else if (live) {
            qs::dyn_tick<EFF>(P, d, rpm, rpm_prev, 0.0, a.substeps, R_last);
        }
        qs::Derived o;
        if (live) { if (a.flags & QS_FLAG_RPY_F32) qs::derive<true>(d, R_last, o); else qs::derive<false>(d, R_last, o); }
        // ---- task ------------------------------------------------------------------------------------------------
        bool env_done = false;
        if (a.task == QS_TASK_HOVER) {
            if (live) {
                const D4 tp = ld256_nc(a.st.target_pos, tbl);
                const qs::TaskTerms tt = qs::hover_terms(P, d, o, tp.x, tp.y, tp.z);
                red_s[2 * t] = tt.reward; red_s[2 * t + 1] = tt.dist; oob_s[t] = tt.out_of_bounds ? 1 : 0;
            }
            __syncthreads();
            if (live && dslot == 0) {
                double rew = 0.0, dist = 0.0; bool oob = false;
                for (int q = 0; q < D; ++q) { rew += red_s[2 * (t + q)]; dist += red_s[2 * (t + q) + 1]; oob |= oob_s[t + q] != 0; }
                const bool term = dist < P.term_dist;
                const bool trunc = oob || ((double)sc / P.pyb_freq > P.episode_len_sec);
                const long long oe = (long long)k * E + e;
                a.io.reward[oe] = (float)rew; a.io.terminated[oe] = term ? 1 : 0; a.io.truncated[oe] = trunc ? 1 : 0;
                if (a.io.done) a.io.done[oe] = (term || trunc) ? 1 : 0;
                done_s[le] = (term || trunc) ? 1 : 0;
            }
            __syncthreads();
            if (live) env_done = done_s[le] != 0;
        } else if (live && dslot == 0) {
            const long long oe = (long long)k * E + e;
            a.io.reward[oe] = -1.0f; a.io.terminated[oe] = 0; a.io.truncated[oe] = 0;
            if (a.io.done) a.io.done[oe] = 0;
        }
        // ---- autoreset, head, bookkeeping ----------------------------------------------------------------------------
        if (live) {
            float* row = base + A + (size_t)t * od;              // my row in the NEXT window
            if ((a.flags & QS_FLAG_AUTORESET_SAME_STEP) && env_done) {
                if (a.flags & QS_FLAG_AUTORESET_CLEARS_HISTORY) for (int q = 12; q < od; ++q) row[q] = 0.f;
                if (a.flags & QS_FLAG_AUTORESET_CLEARS_PID) pst = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                init_drone(a.st, tbl, d);
                const double Rr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                if (a.flags & QS_FLAG_RPY_F32) qs::derive<true>(d, Rr, o); else qs::derive<false>(d, Rr, o);
                rpm[0] = rpm[1] = rpm[2] = rpm[3] = 0.0;
                sc = -a.substeps;
            }
            row[0] = (float)d.px; row[1] = (float)d.py; row[2] = (float)d.pz;
            row[3] = (float)o.roll; row[4] = (float)o.pitch; row[5] = (float)o.yaw;
            row[6] = (float)d.vx; row[7] = (float)d.vy; row[8] = (float)d.vz;
            row[9] = (float)o.ax; row[10] = (float)o.ay; row[11] = (float)o.az;
            sc += a.substeps;
            if (k < T - 1) round_to_planes(d);                // (the final store_drone applies the same rounding once)
            rpm_prev[0] = rpm[0]; rpm_prev[1] = rpm[1]; rpm_prev[2] = rpm[2]; rpm_prev[3] = rpm[3];
        }
        __syncthreads();
        // ---- stream the CTA's rows out: obs[k][c0 .. c0+rows) = window shifted by one action ----------------------------
        base += A;
        {
            float* outp = a.io.obs + ((long long)k * N + c0) * od;
            float* lastp = (k == T - 1 && a.io.obs_last) ? a.io.obs_last + c0 * od : nullptr;
            if (A == 4 && a.stage_mode == 1) {
                // TMA bulk store of the window (see step_kernel); the window is rewritten next tick, so wait until the
                // copy engine has read it
                if (t == 0) {
                    const unsigned bytes = (unsigned)(rows * od * 4);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(outp), "r"(smem_u32(base)), "r"(bytes) : "memory");
                    if (lastp) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(lastp), "r"(smem_u32(base)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            } else if (A == 4) {
                const float4* src = reinterpret_cast<const float4*>(base);
                float4* out = reinterpret_cast<float4*>(outp);
                float4* last = reinterpret_cast<float4*>(lastp);
                const int n4 = rows * (od >> 2), nt = blockDim.x;
                int j = t;
                for (; j + 5 * nt < n4; j += 6 * nt) {
                    const float4 v0 = src[j], v1 = src[j + nt], v2 = src[j + 2 * nt], v3 = src[j + 3 * nt], v4 = src[j + 4 * nt], v5 = src[j + 5 * nt];
                    out[j] = v0; out[j + nt] = v1; out[j + 2 * nt] = v2; out[j + 3 * nt] = v3; out[j + 4 * nt] = v4; out[j + 5 * nt] = v5;
                    if (last) { last[j] = v0; last[j + nt] = v1; last[j + 2 * nt] = v2; last[j + 3 * nt] = v3; last[j + 4 * nt] = v4; last[j + 5 * nt] = v5; }
                }
                for (; j < n4; j += nt) { const float4 v = src[j]; out[j] = v; if (last) last[j] = v; }
            } else {
                for (int j = t; j < rows * od; j += blockDim.x) { const float v = base[j]; outp[j] = v; if (lastp) lastp[j] = v; }
            }
        }
        __syncthreads();
    }
    if (live) {
        store_drone(a.st, N, i, d);
        if (a.st.last_rpm) st256(a.st.last_rpm, i, rpm[0], rpm[1], rpm[2], rpm[3]);
        if (PIDACT) store_pid(a.st.pid, N, i, pst);
        if (dslot == 0) a.st.step_counter[e] = sc;
    }
}

}  // namespace

extern "C" {

int qs_sizeof_rollout_io(void) { return (int)sizeof(QsRolloutIO); }

int qs_rollout_max_ticks(int act_type, int act_buffer_size, int drones_per_env) {
    const int A = act_width(act_type);
    if (A < 0 || act_type == QS_ACT_RAW_RPM || act_buffer_size <= 0 || drones_per_env <= 0 || drones_per_env > kMaxTPB) return 0;
    const size_t span = (size_t)block_size_for(drones_per_env) * (12 + act_buffer_size * A) * 4;
    if (span + 2 * (size_t)A * 4 > kStageLimit) return 0;
    return (int)((kStageLimit - span) / ((size_t)A * 4)) - 1;
}

int qs_rollout(const QsParams* p, const QsState* st, const QsRolloutIO* io, int act_type, int task,
               int n_envs, int drones_per_env, int substeps, unsigned effects, unsigned flags, void* stream) {
    if (!p || !io) return fail(QS_ERR_NULL, "qs_rollout: NULL params/io");
    if (int rc = check_state(st, (flags & QS_FLAG_AUTORESET_SAME_STEP) ? 1 : 0)) return rc;
    if (n_envs <= 0 || drones_per_env <= 0 || substeps <= 0 || io->T <= 0) return fail(QS_ERR_SIZE, "qs_rollout: sizes must be > 0");
    const int A = act_width(act_type);
    if (A < 0 || act_type == QS_ACT_RAW_RPM) return fail(QS_ERR_ENUM, "qs_rollout: bad act_type");
    if (task != QS_TASK_NONE && task != QS_TASK_HOVER) return fail(QS_ERR_ENUM, "qs_rollout: bad task");
    if (effects & ~7u) return fail(QS_ERR_ENUM, "qs_rollout: bad effects");
    if (flags & (QS_FLAG_AUTORESET_NEXT_STEP | QS_FLAG_SKIP_EPILOGUE | QS_FLAG_RPM_FROM_LAST | QS_FLAG_OBS_STATE20))
        return fail(QS_ERR_UNSUPPORTED, "qs_rollout: only SAME_STEP autoreset (or none) is supported");
    if (drones_per_env > kMaxTPB) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: drones_per_env <= 128");
    if (!io->obs_init || !io->obs || !io->reward || !io->terminated || !io->truncated) return fail(QS_ERR_NULL, "qs_rollout: NULL buffer");
    if (io->act_buffer_size <= 0) return fail(QS_ERR_SIZE, "qs_rollout: act_buffer_size must be > 0");
    if (io->T > qs_rollout_max_ticks(act_type, io->act_buffer_size, drones_per_env)) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: T exceeds qs_rollout_max_ticks (split the rollout)");
    if (task == QS_TASK_HOVER && (!st->target_pos || !aligned32(st->target_pos))) return fail(QS_ERR_NULL, "qs_rollout: target_pos NULL/misaligned");
    const bool pid_act = act_type == QS_ACT_PID || act_type == QS_ACT_VEL || act_type == QS_ACT_ONE_D_PID;
    if (pid_act && !st->pid) return fail(QS_ERR_NULL, "qs_rollout: PID action type needs QsState.pid");
    if ((effects & QS_EFFECT_DRAG) && !st->last_rpm) return fail(QS_ERR_NULL, "qs_rollout: DRAG needs QsState.last_rpm");
    if (A == 4 && ((io->actions && !aligned16(io->actions)) || (io->actions_out && !aligned16(io->actions_out)) || !aligned16(io->obs) || (io->obs_last && !aligned16(io->obs_last))))
        return fail(QS_ERR_ALIGN, "qs_rollout: [N][4]-wide buffers must be 16-byte aligned");
    RolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.P = *p; a.st = *st; a.io = *io;
    a.act_type = act_type; a.task = task; a.n_envs = n_envs; a.D = drones_per_env; a.substeps = substeps;
    if ((long long)n_envs * drones_per_env > 0x7fffffffLL) return fail(QS_ERR_SIZE, "qs_rollout: n_envs * drones_per_env exceeds 2^31-1");
    a.N = n_envs * drones_per_env; a.A = A; a.obs_dim = 12 + io->act_buffer_size * A;
    a.cap = cta_capacity(a.N, drones_per_env, true);
    a.tpb = block_size_for(drones_per_env, a.cap);
    a.effects = effects; a.flags = flags;
    {
        const size_t row_bytes = (size_t)a.obs_dim * 4, span = row_bytes * a.tpb;
        const bool aligned = aligned16(io->obs_init) && (span % 16 == 0) && ((row_bytes * ((size_t)a.N % a.tpb)) % 16 == 0);
        a.stage_mode = (aligned && A == 4) ? 1 : 2;
    }
    const int blocks = (int)((a.N + a.tpb - 1) / a.tpb);
    int threads = ((a.tpb + 31) / 32) * 32;
    size_t sm = smem_fixed(a.cap) + (size_t)a.tpb * a.obs_dim * 4 + (size_t)(io->T + 1) * A * 4 + 32;
    cudaStream_t s = (cudaStream_t)stream;
    if (io->policy) {
        const QsPolicy& q = *io->policy;
        if (pid_act || (effects & 7u) || a.cap > 64) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: the on-device policy supports RPM / ONE_D_RPM actions, no DYN+ effects, drones_per_env <= 64");
        if (io->actions) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: pass either actions or a policy");
        if (!q.w1 || !q.b1 || !q.w2 || !q.b2 || !q.w3 || !q.b3 || !q.log_std) return fail(QS_ERR_NULL, "qs_rollout: policy weights are NULL");
        if (q.nt3 != 1 && q.nt3 != 2 && q.nt3 != 4) return fail(QS_ERR_SIZE, "qs_rollout: policy nt3 (padded output tiles of 8) must be 1, 2 or 4");
        if (q.out_dim > 8 * q.nt3) return fail(QS_ERR_SIZE, "qs_rollout: policy out_dim exceeds the padded output width");
        if (q.in_dim != drones_per_env * a.obs_dim || q.out_dim != drones_per_env * A) return fail(QS_ERR_SIZE, "qs_rollout: policy in_dim/out_dim must be D*obs_dim / D*A");
        if (q.vw1 && (!q.vb1 || !q.vw2 || !q.vb2 || !q.vw3 || !q.vb3)) return fail(QS_ERR_NULL, "qs_rollout: incomplete critic");
        if (q.values && !q.vw1) return fail(QS_ERR_NULL, "qs_rollout: values requested without a critic");
        if (!aligned16(q.w1) || !aligned16(q.w2) || !aligned16(q.w3) || (q.vw1 && (!aligned16(q.vw1) || !aligned16(q.vw2) || !aligned16(q.vw3))))
            return fail(QS_ERR_ALIGN, "qs_rollout: policy weight arrays must be 16-byte aligned");
        a.pol = q;
        threads = 64;                                            // two warps: 64 drones, 32 hidden units each in the MLP
        const int n_av_max = 64 / drones_per_env;
        sm = 1280 + (size_t)a.tpb * a.obs_dim * 4 + (size_t)(io->T + 1) * A * 4 + 32
           + (size_t)(n_av_max * 8 * q.nt3 + n_av_max * 8 + 64) * 4 + 16;      // compact fixed part, window, means, values, log-prob terms
        if (sm > 200 * 1024) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: policy + window exceed shared memory");
        if (sm > 48 * 1024) cudaFuncSetAttribute(rollout_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        {   // shared-memory carve-out: just enough for the 7 resident CTAs, so that the weights find the rest of the 256 KB as L1
            const size_t need = 7 * (sm + 1024);
            int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
            cudaFuncSetAttribute(rollout_kernel<0, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
        }
        rollout_kernel<0, false, true><<<blocks, threads, sm, s>>>(a);
        const cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? 0 : cuda_fail(e, "qs_rollout (policy) launch");
    }
#define QS_RCASE(E)                                                                                                   \
    case E: {                                                                                                         \
        if (pid_act) {                                                                                                \
            if (sm > 48 * 1024) cudaFuncSetAttribute(rollout_kernel<E, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStepSmemFixed + kStageLimit + 32)); \
            rollout_kernel<E, true, false><<<blocks, threads, sm, s>>>(a);                                            \
        } else {                                                                                                      \
            if (sm > 48 * 1024) cudaFuncSetAttribute(rollout_kernel<E, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStepSmemFixed + kStageLimit + 32)); \
            rollout_kernel<E, false, false><<<blocks, threads, sm, s>>>(a);                                           \
        }                                                                                                             \
    } break;
    switch (effects & 7u) { QS_RCASE(0) QS_RCASE(1) QS_RCASE(2) QS_RCASE(3) QS_RCASE(4) QS_RCASE(5) QS_RCASE(6) QS_RCASE(7) }
#undef QS_RCASE
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_rollout launch");
}

}  // extern "C"
