// rollout.cu -- qs_rollout: T fused control ticks per launch.
#include "qs_common.cuh"

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
    const float* aw[6];  // actor: {w1 hi, w1 lo, w2 hi, w2 lo, w3 hi, w3 lo}
    const float* cw[6];  // critic
    int act_type, task, n_envs, D, substeps, N, A, obs_dim, tpb;
    unsigned effects, flags;
    int stage_mode, cap;
};

// ---- on-device policy: SB3-MlpPolicy-shaped MLP on the tensor cores, fp32-accurate ------------------------------------------------
// Y[32 aviaries][64 units] = X[32][K] W[K][64] per layer and warp, as mma.sync.m16n8k8 TF32 tiles with the 3xTF32 split:
// x = x_hi + x_lo, w = w_hi + w_lo (each part exactly representable in TF32), x w ~ x_hi w_hi + x_hi w_lo + x_lo w_hi with fp32
// accumulation -- relative error ~2^-21 per product, i.e. fp32-level (a plain TF32/BF16 mma has 2^-11 / 2^-8 and fails the 1e-5
// parity with the fp32 torch network).  The weights are split once on the host (MlpPolicy); the activations are split per
// fragment load.  K = 144 / 64 and M = 32 rows per warp are far below a tcgen05 tile (M = 128, operands through shared-memory
// descriptors and TMEM): at 1.8 GFLOP per tick the legacy mma.sync path is already not the bottleneck.
constexpr int kHid = 64, kHidStride = 68;      // hidden rows padded to 68 floats: conflict-free fragment loads (4 g + t distinct banks)

__device__ __forceinline__ void tf32_split(float x, unsigned& hi, unsigned& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float c[4], const unsigned a[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One layer for MT m-tiles of 16 aviaries (rows row0 .. row0 + 16 MT) and NT n-tiles of 8 units.  x rows: K floats at stride
// x_stride in shared memory; Whi / Wlo: [Kpad][8 NT] row-major TF32 halves in global memory (through L1), Kpad = K rounded up to 8
// with zero rows; bias [8 NT].  Result (tanh or identity) to y_s[row][y_stride] (may alias x_s: all reads precede the writes).
template <int MT, int NT, bool TANH>
__device__ __forceinline__ void mma_layer(const float* x_s, int x_stride, int K, const float* __restrict__ Whi, const float* __restrict__ Wlo,
                                          const float* __restrict__ bias, float* y_s, int y_stride, int row0, int n_rows, int lane) {
    const int g = lane >> 2, t = lane & 3;
    float c[MT][NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const float b0 = __ldg(bias + 8 * n + 2 * t), b1 = __ldg(bias + 8 * n + 2 * t + 1);
#pragma unroll
        for (int m = 0; m < MT; ++m) { c[m][n][0] = b0; c[m][n][1] = b1; c[m][n][2] = b0; c[m][n][3] = b1; }
    }
    const int ldw = 8 * NT;
    for (int k0 = 0; k0 < K; k0 += 8) {
        unsigned ahi[MT][4], alo[MT][4];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int r0 = row0 + 16 * m + g, r1 = r0 + 8;
            const bool k_lo = k0 + t < K, k_hi = k0 + t + 4 < K;
            const float a0 = (r0 < n_rows && k_lo) ? x_s[(size_t)r0 * x_stride + k0 + t] : 0.f;
            const float a1 = (r1 < n_rows && k_lo) ? x_s[(size_t)r1 * x_stride + k0 + t] : 0.f;
            const float a2 = (r0 < n_rows && k_hi) ? x_s[(size_t)r0 * x_stride + k0 + t + 4] : 0.f;
            const float a3 = (r1 < n_rows && k_hi) ? x_s[(size_t)r1 * x_stride + k0 + t + 4] : 0.f;
            tf32_split(a0, ahi[m][0], alo[m][0]); tf32_split(a1, ahi[m][1], alo[m][1]);
            tf32_split(a2, ahi[m][2], alo[m][2]); tf32_split(a3, ahi[m][3], alo[m][3]);
        }
        const float* wh = Whi + (size_t)(k0 + t) * ldw + g;
        const float* wl = Wlo + (size_t)(k0 + t) * ldw + g;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const unsigned bh0 = __float_as_uint(__ldg(wh + 8 * n)), bh1 = __float_as_uint(__ldg(wh + 4 * ldw + 8 * n));
            const unsigned bl0 = __float_as_uint(__ldg(wl + 8 * n)), bl1 = __float_as_uint(__ldg(wl + 4 * ldw + 8 * n));
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                mma_tf32(c[m][n], alo[m], bh0, bh1);          // small terms first
                mma_tf32(c[m][n], ahi[m], bl0, bl1);
                mma_tf32(c[m][n], ahi[m], bh0, bh1);
            }
        }
    }
    __syncwarp();                                            // in-place layers: every read of x precedes the first write of y
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const int r0 = row0 + 16 * m + g, r1 = r0 + 8;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            float v0 = c[m][n][0], v1 = c[m][n][1], v2 = c[m][n][2], v3 = c[m][n][3];
            if (TANH) { v0 = tanhf(v0); v1 = tanhf(v1); v2 = tanhf(v2); v3 = tanhf(v3); }
            if (r0 < n_rows) *reinterpret_cast<float2*>(y_s + (size_t)r0 * y_stride + 8 * n + 2 * t) = make_float2(v0, v1);
            if (r1 < n_rows) *reinterpret_cast<float2*>(y_s + (size_t)r1 * y_stride + 8 * n + 2 * t) = make_float2(v2, v3);
        }
    }
    __syncwarp();
}

// One network (in -> 64 tanh -> 64 tanh -> 8 NT3 outputs, of which n_out are real) for MT m-tiles of aviaries starting at row0.
// out_s rows have stride 8 NT3 floats (padded outputs).
template <int MT>
__device__ __forceinline__ void mma_net(const float* x_s, int in_dim, const float* const* W /* hi/lo of the three layers */, const float* b1,
                                        const float* b2, const float* b3, int nt3, float* h_s, float* out_s, int row0, int n_rows, int lane) {
    mma_layer<MT, 8, true>(x_s, in_dim, in_dim, W[0], W[1], b1, h_s, kHidStride, row0, n_rows, lane);
    mma_layer<MT, 8, true>(h_s, kHidStride, kHid, W[2], W[3], b2, h_s, kHidStride, row0, n_rows, lane);
    if (nt3 == 1) mma_layer<MT, 1, false>(h_s, kHidStride, kHid, W[4], W[5], b3, out_s, 8, row0, n_rows, lane);
    else if (nt3 == 2) mma_layer<MT, 2, false>(h_s, kHidStride, kHid, W[4], W[5], b3, out_s, 16, row0, n_rows, lane);
    else mma_layer<MT, 4, false>(h_s, kHidStride, kHid, W[4], W[5], b3, out_s, 32, row0, n_rows, lane);
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
    // POLICY: [2 nets][32 aviaries][68] hidden activations, [32][out_dim] means, [32] values, [tpb] log-prob terms, after the window
    // POLICY scratch after the window: per warp one tile of 16 hidden rows [16][68]; padded action means [n_av][8 nt3]; padded
    // values [n_av][8]; log-prob terms [64]
    float* pol_s = stage_s + ((((size_t)tpb * od + (size_t)(T + 1) * A) + 3) & ~(size_t)3);
    float* mean_s = pol_s + 2 * 16 * kHidStride;
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
            const int n_av = rows / D, warp = t >> 5, lane = t & 31;
            const int ost = 8 * a.pol.nt3;                        // padded width of the output rows in mean_s
            // 16 aviaries (one m-tile) per warp and call: warp 0 runs the actor, warp 1 the critic, over the same tiles; without a
            // critic the two warps alternate tiles of the actor.  Hidden rows: one [16][68] tile per warp, overwritten in place.
            float* hid = pol_s + (size_t)warp * 16 * kHidStride;
            for (int r0 = 0; r0 < n_av; r0 += 16) {
                const float* x0 = base + (size_t)r0 * a.pol.in_dim;
                const int na = n_av - r0;                         // rows of this tile that exist (mma_layer clips at 16)
                if (a.pol.vw1) {
                    if (warp == 0) mma_net<1>(x0, a.pol.in_dim, a.aw, a.pol.b1, a.pol.b2, a.pol.b3, a.pol.nt3, hid, mean_s + (size_t)r0 * ost, 0, na, lane);
                    else if (warp == 1) mma_net<1>(x0, a.pol.in_dim, a.cw, a.pol.vb1, a.pol.vb2, a.pol.vb3, 1, hid, val_s + (size_t)r0 * 8, 0, na, lane);
                } else if (warp == ((r0 >> 4) & 1)) {
                    mma_net<1>(x0, a.pol.in_dim, a.aw, a.pol.b1, a.pol.b2, a.pol.b3, a.pol.nt3, hid, mean_s + (size_t)r0 * ost, 0, na, lane);
                }
            }
            __syncthreads();
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
        if (!q.w1 || !q.w1_lo || !q.b1 || !q.w2 || !q.w2_lo || !q.b2 || !q.w3 || !q.w3_lo || !q.b3 || !q.log_std) return fail(QS_ERR_NULL, "qs_rollout: policy weights are NULL");
        if (q.nt3 != 1 && q.nt3 != 2 && q.nt3 != 4) return fail(QS_ERR_SIZE, "qs_rollout: policy nt3 (padded output tiles of 8) must be 1, 2 or 4");
        if (q.out_dim > 8 * q.nt3) return fail(QS_ERR_SIZE, "qs_rollout: policy out_dim exceeds the padded output width");
        if (q.in_dim != drones_per_env * a.obs_dim || q.out_dim != drones_per_env * A) return fail(QS_ERR_SIZE, "qs_rollout: policy in_dim/out_dim must be D*obs_dim / D*A");
        if (q.vw1 && (!q.vw1_lo || !q.vb1 || !q.vw2 || !q.vw2_lo || !q.vb2 || !q.vw3 || !q.vw3_lo || !q.vb3)) return fail(QS_ERR_NULL, "qs_rollout: incomplete critic");
        if (q.values && !q.vw1) return fail(QS_ERR_NULL, "qs_rollout: values requested without a critic");
        a.pol = q;
        a.aw[0] = q.w1; a.aw[1] = q.w1_lo; a.aw[2] = q.w2; a.aw[3] = q.w2_lo; a.aw[4] = q.w3; a.aw[5] = q.w3_lo;
        a.cw[0] = q.vw1; a.cw[1] = q.vw1_lo; a.cw[2] = q.vw2; a.cw[3] = q.vw2_lo; a.cw[4] = q.vw3; a.cw[5] = q.vw3_lo;
        threads = 64;                                            // warp 0: actor, warp 1: critic
        const int n_av_max = 64 / drones_per_env;
        sm = 1280 + (size_t)a.tpb * a.obs_dim * 4 + (size_t)(io->T + 1) * A * 4 + 32
           + (size_t)(2 * 16 * kHidStride + n_av_max * 8 * q.nt3 + n_av_max * 8 + 64) * 4 + 16;      // compact fixed part, window, hidden tiles, means, values, log-prob terms
        if (sm > 200 * 1024) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: policy + window exceed shared memory");
        if (sm > 48 * 1024) cudaFuncSetAttribute(rollout_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
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
