// K6: Ising-model tabular mean-field Q-learning, one fused sweep per launch.
//
// Reference (pure Python): main_MFQ_Ising.py:105-134 (loop body), :55-67 (boltzman_explore),
// examples/ising_model/Ising.py:7-58 (4-neighbour torus mask), :101-111 (reward), :113-118 (observation),
// multiagent/core.py:99-125 (spin <- action), multiagent/environment.py:49-78.
// Per site i of an L x L torus, per step:
//     s   = number of up neighbours in the OLD lattice            (obs of the previous step)
//     p_a = exp(Q[i,s,a] / T) / sum_a' exp(Q[i,s,a'] / T);  a = [u >= p_0]   (np.random.choice(2, 1, p))
//     spin_i <- a                                                  (all sites, synchronously)
//     r   = 0.5 * sigma_i * sum_nbr sigma_j on the NEW lattice,  sigma = 2 spin - 1
//     Q[i,s,a] <- Q[i,s,a] + lr * (r - Q[i,s,a])                   (sites in the act group)
//
// Layout in HBM:  spins int8 [B][L][L];  Q  T [B][5][L*L][2]  (one plane per neighbour count s, the action
// pair of a site adjacent: a site reads its pair with ONE 8-byte load and rewrites one half of it, and a
// warp walking a lattice row touches consecutive pairs of one plane whenever neighbouring sites share s.
// Measured on B200 against the [5][2][N] split-plane layout in the disordered phase: fewer partially
// used 32-byte sectors; an AoS [site][5][2] row (40 B) would cost a full sector read + write per site).
//
// Mapping: one CTA per lattice, one thread per lattice COLUMN, warps = ceil(L / 32).  Both the old and
// the new lattice live in shared memory bit-packed (__ballot_sync packs a warp's 32 columns into one
// word), so a 256 x 256 lattice costs 2 x 8 KB and eight CTAs fit per SM.  Rows are swept in bands of
// RB rows: the band's Q pairs are loaded with 2*RB independent loads per thread in flight, actions are
// drawn and the band's new bits published, and after one barrier the rows whose three new neighbour
// rows are known are finalised from registers (reward, Q update) -- Q is read once and written once per
// site.  Row 0 waits in registers for row L-1 (torus wrap).  Algorithmic traffic per site-step:
// 2 x 4 (Q read) + 4 (Q write) + 1 + 1 (spin) = 14 B in fp32.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mfmarl_batched.h"
#include "rng.cuh"
#include "engine.h"

namespace mfmarl {

constexpr int kIsingRB = 4;   // rows per band == uniforms per Philox call

template <typename T>
struct IsingArgs {
    int B, L;
    int8_t *spins;          // [B][L][L] in/out
    T *Q;                   // [B][5][L*L][2] in/out
    T temperature, lr;
    const T *u;             // optional injected uniforms [B][L*L] (test hook); else Philox
    const uint8_t *mask;    // optional act-group mask [B][L*L]; null = every site updates Q
    uint32_t seed, lattice_base, step;
    int32_t *n_up;          // [B] out: up spins after the sweep
    T *reward_sum;          // [B] out: sum_i r_i
    T *mse;                 // [B] out: sum over updated sites (Q_new - target)^2 / N   (main_MFQ_Ising.py:127-134)
};

template <typename T> struct Pair;
template <> struct Pair<float> { typedef float2 type; };
template <> struct Pair<double> { typedef double2 type; };

template <typename T> __device__ __forceinline__ T uniform_from_bits(uint32_t hi, uint32_t lo);
template <> __device__ __forceinline__ float uniform_from_bits<float>(uint32_t hi, uint32_t) {
    return (float)(hi >> 8) * (1.0f / 16777216.0f);                       // 24 bits, [0, 1)
}
template <> __device__ __forceinline__ double uniform_from_bits<double>(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);   // 53 bits
}

// p_0 after numpy's own renormalisation: np.random.choice builds cdf = cumsum(p); cdf /= cdf[-1] and
// returns searchsorted(cdf, u, 'right')  =>  a = [u >= cdf_0].
__device__ __forceinline__ double action_threshold(double q0, double q1, double T) {
    const double e0 = exp(q0 / T), e1 = exp(q1 / T);
    const double denom = 0.0 + e0 + e1;                                   // denom = 0; denom += val (x2)
    const double p0 = e0 / denom, p1 = e1 / denom;
    return p0 / (p0 + p1);
}
__device__ __forceinline__ float action_threshold(float q0, float q1, float T) {
    // same quantity, written so that exp cannot overflow in fp32 at small T: e0/(e0+e1) = 1/(1+exp((q1-q0)/T)).
    // ex2.approx + rcp.approx: |error| of the threshold < 4e-6 for |q| <= 2, T >= 0.25 (2 + 1.16|x| ulp of exp),
    // i.e. an action can differ from the fp64 reference only for draws that close to the threshold -- the same
    // order as the fp32 rounding of u itself (2^-24).  The fp64 kernel mode evaluates the reference's formula.
    return __frcp_rn(1.0f + __expf(__fdividef(q1 - q0, T)));
}

__device__ __forceinline__ int bit_at(const uint32_t *rows, int wpr, int r, int x) {
    return (rows[r * wpr + (x >> 5)] >> (x & 31)) & 1;
}

template <typename T>
__global__ void __launch_bounds__(sizeof(T) == 8 ? 512 : 1024) k_ising(const IsingArgs<T> A) {
    extern __shared__ __align__(16) uint32_t s_bits[];
    const int L = A.L, N = L * L, wpr = (L + 31) >> 5;
    uint32_t *s_old = s_bits, *s_new = s_bits + L * wpr;
    __shared__ int s_nup;
    __shared__ T s_rsum, s_mse;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int x = tid;                       // this thread's column
    const bool active = x < L;
    const size_t lbase = (size_t)b * N;
    int8_t *spins = A.spins + lbase;
    T *Q = A.Q + lbase * 10;
    const int xl = x == 0 ? L - 1 : x - 1, xr = x == L - 1 ? 0 : x + 1;

    if (tid == 0) { s_nup = 0; s_rsum = (T)0; s_mse = (T)0; }
    // ---- old lattice -> bits ----
#pragma unroll 8
    for (int r = 0; r < L; r++) {
        const int v = active ? spins[r * L + x] : 0;
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, v != 0);
        if (lane == 0) s_old[r * wpr + w] = word;
    }
    __syncthreads();

    T keep_q[kIsingRB]; int keep_sa[kIsingRB];     // this band: selected Q value, s | a << 3
    T prev_q = (T)0; int prev_sa = 0;              // last row of the previous band
    T row0_q = (T)0; int row0_sa = 0;              // row 0, finalised last (needs row L-1)
    int nup = 0; T rsum = (T)0, mse = (T)0;

    auto finalize = [&](int r, T qsel, int sa) {
        const int s = sa & 7, a = sa >> 3, i = r * L + x;
        const int ru = r == 0 ? L - 1 : r - 1, rd = r == L - 1 ? 0 : r + 1;
        const int ups = bit_at(s_new, wpr, ru, x) + bit_at(s_new, wpr, rd, x) + bit_at(s_new, wpr, r, xl) +
                        bit_at(s_new, wpr, r, xr);
        const T reward = (T)0.5 * (T)(2 * a - 1) * (T)(2 * ups - 4);      // Ising.py:101-111
        spins[i] = (int8_t)a;
        nup += a; rsum += reward;
        if (A.mask == nullptr || A.mask[lbase + i]) {
            const T qn = qsel + A.lr * (reward - qsel);                    // main_MFQ_Ising.py:129-131
            Q[((size_t)s * N + i) * 2 + a] = qn;
            const T d = qn - (T)((2 - s) * (1 - 2 * a));                   // reward_target[s][a], :77-81
            mse += d * d;
        }
    };

    for (int r0 = 0; r0 < L; r0 += kIsingRB) {
        const int nrows = min(kIsingRB, L - r0);
        T q0[kIsingRB], q1[kIsingRB]; int sv[kIsingRB];
        // ---- neighbour counts on the OLD lattice and the band's Q pairs: 2*RB loads in flight ----
#pragma unroll
        for (int j = 0; j < kIsingRB; j++) {
            sv[j] = 0; q0[j] = (T)0; q1[j] = (T)0;
            if (j < nrows && active) {
                const int r = r0 + j, ru = r == 0 ? L - 1 : r - 1, rd = r == L - 1 ? 0 : r + 1;
                sv[j] = bit_at(s_old, wpr, ru, x) + bit_at(s_old, wpr, rd, x) + bit_at(s_old, wpr, r, xl) +
                        bit_at(s_old, wpr, r, xr);                         // Ising.py:113-118 + count_nonzero
                const size_t i = (size_t)r * L + x;
                const typename Pair<T>::type pr = *(const typename Pair<T>::type *)(Q + ((size_t)sv[j] * N + i) * 2);
                q0[j] = pr.x; q1[j] = pr.y;
            }
        }
        // ---- uniforms: one Philox call per thread per band, keyed (seed, lattice) x (column, band, step) ----
        T uu[kIsingRB];
        if (A.u != nullptr) {
#pragma unroll
            for (int j = 0; j < kIsingRB; j++)
                uu[j] = (j < nrows && active) ? A.u[lbase + (size_t)(r0 + j) * L + x] : (T)0;
        } else if (sizeof(T) == 4) {
            const uint4 rnd = philox4x32_10(make_uint4((uint32_t)x, (uint32_t)(r0 / kIsingRB), A.step, 0u),
                                            make_uint2(A.seed, A.lattice_base + (uint32_t)b));
            uu[0] = uniform_from_bits<T>(rnd.x, 0); uu[1] = uniform_from_bits<T>(rnd.y, 0);
            uu[2] = uniform_from_bits<T>(rnd.z, 0); uu[3] = uniform_from_bits<T>(rnd.w, 0);
        } else {
            const uint2 key = make_uint2(A.seed, A.lattice_base + (uint32_t)b);
            const uint4 r1 = philox4x32_10(make_uint4((uint32_t)x, (uint32_t)(r0 / kIsingRB), A.step, 0u), key);
            const uint4 r2 = philox4x32_10(make_uint4((uint32_t)x, (uint32_t)(r0 / kIsingRB), A.step, 1u), key);
            uu[0] = uniform_from_bits<T>(r1.x, r1.y); uu[1] = uniform_from_bits<T>(r1.z, r1.w);
            uu[2] = uniform_from_bits<T>(r2.x, r2.y); uu[3] = uniform_from_bits<T>(r2.z, r2.w);
        }
        // ---- Boltzmann draw, publish the band's new bits ----
#pragma unroll
        for (int j = 0; j < kIsingRB; j++) {
            int a = 0;
            if (j < nrows && active) a = uu[j] >= action_threshold(q0[j], q1[j], A.temperature) ? 1 : 0;
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, a != 0);
            if (j < nrows && lane == 0) s_new[(r0 + j) * wpr + w] = word;
            keep_q[j] = a ? q1[j] : q0[j];
            keep_sa[j] = sv[j] | (a << 3);
        }
        __syncthreads();
        // ---- finalise every row whose three new neighbour rows are now known ----
        if (active) {
            if (r0 > 0) {
                if (r0 - 1 > 0) finalize(r0 - 1, prev_q, prev_sa);
            }
#pragma unroll
            for (int j = 0; j < kIsingRB; j++) {
                const int r = r0 + j;
                if (j < nrows - 1) {
                    if (r == 0) { row0_q = keep_q[j]; row0_sa = keep_sa[j]; }
                    else finalize(r, keep_q[j], keep_sa[j]);
                }
            }
        }
        // remember the band's last row for the next iteration
#pragma unroll
        for (int j = 0; j < kIsingRB; j++)
            if (j == nrows - 1) { prev_q = keep_q[j]; prev_sa = keep_sa[j]; }
    }
    if (active) {
        if (L > 1) finalize(L - 1, prev_q, prev_sa);   // needs row 0, published in the first band
        finalize(0, row0_q, row0_sa);                  // needs row L-1, published in the last band
    }

    // ---- per-lattice statistics: up count, reward sum, mse ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nup += __shfl_xor_sync(0xFFFFFFFFu, nup, o);
        rsum += __shfl_xor_sync(0xFFFFFFFFu, rsum, o);
        mse += __shfl_xor_sync(0xFFFFFFFFu, mse, o);
    }
    if (lane == 0) { atomicAdd(&s_nup, nup); atomicAdd(&s_rsum, rsum); atomicAdd(&s_mse, mse); }
    __syncthreads();
    if (tid == 0) {
        if (A.n_up) A.n_up[b] = s_nup;
        if (A.reward_sum) A.reward_sum[b] = s_rsum;
        if (A.mse) A.mse[b] = s_mse / (T)N;
    }
}

template <typename T>
static void launch_ising(const IsingArgs<T> &A, cudaStream_t st) {
    const int L = A.L, wpr = (L + 31) >> 5;
    if (L < 3 || L > 1024) throw Fatal("ising: lattice side must be in [3, 1024]");
    if ((size_t)2 * L * wpr * sizeof(uint32_t) > 200 * 1024) throw Fatal("ising: lattice too large for the shared-memory bit planes");
    if (A.B < 1) throw Fatal("ising: need at least one lattice");
    if (sizeof(T) == 8 && L > 512) throw Fatal("ising: fp64 mode supports lattice sides up to 512");
    const int threads = wpr * 32;
    const size_t smem = (size_t)2 * L * wpr * sizeof(uint32_t);
    if (smem > 48 * 1024)
        MF_CUDA(cudaFuncSetAttribute(k_ising<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ising<T><<<A.B, threads, smem, st>>>(A);
    MF_CUDA(cudaGetLastError());
}


// ----------------------------------------------------------------------------------------------
// K6r: RESIDENT variant -- K sweeps per launch with the Q table in shared memory.
//
// The streaming kernel above moves every site's Q pair through HBM once per sweep (14 B/site-step
// algorithmic, 45-60 B real in the disordered phase because the plane index s is data dependent and
// 32-byte sectors are only partly used).  Here a lattice is cut into C horizontal strips, one CTA each,
// the C CTAs forming one thread-block cluster; a CTA keeps its strip of Q ([5][rows*L][2], up to 160 KB) and
// both bit-packed spin buffers in shared memory for all K sweeps and reads the neighbour strips' boundary
// rows through distributed shared memory.  One cluster barrier per sweep suffices: a sweep reads the "old"
// buffer (own + halo) and writes only its own "new" buffer, the buffers swap every sweep, and a CTA can
// only be one barrier ahead of its neighbours.  HBM traffic drops to (80 B load + 80 B store) / K per
// site-step.  Draws use the same Philox keys as the streaming kernel -- (seed, lattice) x (column, global
// row band, step) -- so K resident sweeps equal K streaming launches bit for bit.
// C = 1 covers lattices whose whole Q fits one CTA (L <= 64 in fp32; the reference's 20 x 20 case);
// 256 x 256 uses C = 16 (16 rows per CTA, 1024 threads = 256 columns x 4 bands).
// ----------------------------------------------------------------------------------------------
}  // namespace mfmarl
#include <cooperative_groups.h>
namespace mfmarl {
namespace cg = cooperative_groups;

template <typename T>
struct IsingRunArgs {
    int B, L, K, rows_per;         // lattices, side, sweeps, rows per CTA (= L / cluster size)
    int8_t *spins; T *Q;
    const T *temperatures;          // [K] device
    T lr;
    const T *u;                     // optional injected uniforms [K][B][L*L]
    uint32_t seed, lattice_base, step0;
    int32_t *n_up;                  // [K][B] out, must be zeroed by the caller
    T *reward_sum;                  // [K][B] out or null, zeroed by the caller
};

template <typename T>
__global__ void __launch_bounds__(1024, 1) k_ising_resident(const IsingRunArgs<T> A) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int L = A.L, N = L * L, wpr = (L + 31) >> 5, LP = wpr * 32, rows = A.rows_per;
    const int b = blockIdx.x / C, row0 = rank * rows;
    const int strip = rows * L;                          // sites in this CTA's strip

    T *s_q = (T *)s_raw;                                 // [5][strip][2]
    uint32_t *s_bits = (uint32_t *)(s_raw + (size_t)5 * strip * 2 * sizeof(T));   // [2][rows][wpr]
    __shared__ int s_red_i;
    __shared__ T s_red_r;

    const int tid = threadIdx.x, lane = tid & 31;
    const int x = tid % LP, band = tid / LP, w = x >> 5;  // column, band (4 rows) inside the strip
    const bool active = x < L;
    const int xl = x == 0 ? L - 1 : x - 1, xr = x == L - 1 ? 0 : x + 1;
    const size_t lbase = (size_t)b * N;

    // ---- load: Q strip (5 contiguous plane slices) and spins -> bits ----
    for (int sp = 0; sp < 5; sp++) {
        const T *src = A.Q + (lbase * 5 + (size_t)sp * N + (size_t)row0 * L) * 2;
        T *dst = s_q + (size_t)sp * strip * 2;
        for (int i = tid; i < strip * 2; i += blockDim.x) dst[i] = src[i];
    }
#pragma unroll
    for (int j = 0; j < kIsingRB; j++) {
        const int r = band * kIsingRB + j;
        int v = 0;
        if (r < rows && active) v = A.spins[lbase + (size_t)(row0 + r) * L + x];
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, v != 0);
        if (r < rows && lane == 0) s_bits[r * wpr + w] = word;
    }
    cluster.sync();

    const int up_rank = (rank + C - 1) % C, dn_rank = (rank + 1) % C;
    int cur = 0;                                          // which bit buffer holds the current lattice
    for (int k = 0; k < A.K; k++, cur ^= 1) {
        const uint32_t *old_own = s_bits + cur * rows * wpr;
        uint32_t *new_own = s_bits + (cur ^ 1) * rows * wpr;
        const uint32_t *old_up = cluster.map_shared_rank(s_bits, up_rank) + cur * rows * wpr + (rows - 1) * wpr;
        const uint32_t *old_dn = cluster.map_shared_rank(s_bits, dn_rank) + cur * rows * wpr;
        const T temperature = A.temperatures[k];
        if (tid == 0) { s_red_i = 0; s_red_r = (T)0; }

        // ---- phase 1: neighbour counts on the old lattice, Boltzmann draw, publish new bits ----
        T keep_q[kIsingRB]; int keep_sa[kIsingRB];
        T uu[kIsingRB];
        const int gband = (row0 >> 2) + band;             // global band index: same Philox counter as k_ising
        if (A.u != nullptr) {
#pragma unroll
            for (int j = 0; j < kIsingRB; j++) {
                const int r = band * kIsingRB + j;
                uu[j] = (r < rows && active) ? A.u[((size_t)k * A.B + b) * N + (size_t)(row0 + r) * L + x] : (T)0;
            }
        } else if (sizeof(T) == 4) {
            const uint4 rnd = philox4x32_10(make_uint4((uint32_t)x, (uint32_t)gband, A.step0 + (uint32_t)k, 0u),
                                            make_uint2(A.seed, A.lattice_base + (uint32_t)b));
            uu[0] = uniform_from_bits<T>(rnd.x, 0); uu[1] = uniform_from_bits<T>(rnd.y, 0);
            uu[2] = uniform_from_bits<T>(rnd.z, 0); uu[3] = uniform_from_bits<T>(rnd.w, 0);
        } else {
            const uint2 key = make_uint2(A.seed, A.lattice_base + (uint32_t)b);
            const uint4 r1 = philox4x32_10(make_uint4((uint32_t)x, (uint32_t)gband, A.step0 + (uint32_t)k, 0u), key);
            const uint4 r2 = philox4x32_10(make_uint4((uint32_t)x, (uint32_t)gband, A.step0 + (uint32_t)k, 1u), key);
            uu[0] = uniform_from_bits<T>(r1.x, r1.y); uu[1] = uniform_from_bits<T>(r1.z, r1.w);
            uu[2] = uniform_from_bits<T>(r2.x, r2.y); uu[3] = uniform_from_bits<T>(r2.z, r2.w);
        }
#pragma unroll
        for (int j = 0; j < kIsingRB; j++) {
            const int r = band * kIsingRB + j;
            int a = 0, sv = 0; T q0 = (T)0, q1 = (T)0;
            if (r < rows && active) {
                const int up = r == 0 ? (int)((old_up[x >> 5] >> (x & 31)) & 1u) : bit_at(old_own, wpr, r - 1, x);
                const int dn = r == rows - 1 ? (int)((old_dn[x >> 5] >> (x & 31)) & 1u) : bit_at(old_own, wpr, r + 1, x);
                sv = up + dn + bit_at(old_own, wpr, r, xl) + bit_at(old_own, wpr, r, xr);
                const typename Pair<T>::type pr = *(const typename Pair<T>::type *)(s_q + ((size_t)sv * strip + r * L + x) * 2);
                q0 = pr.x; q1 = pr.y;
                a = uu[j] >= action_threshold(q0, q1, temperature) ? 1 : 0;
            }
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, a != 0);
            if (r < rows && lane == 0) new_own[r * wpr + w] = word;
            keep_q[j] = a ? q1 : q0;
            keep_sa[j] = sv | (a << 3);
        }
        cluster.sync();   // every strip's new bits are published (and the reduction cells are reset)

        // ---- phase 2: reward on the new lattice, Q update in shared memory ----
        const uint32_t *new_up = cluster.map_shared_rank(s_bits, up_rank) + (cur ^ 1) * rows * wpr + (rows - 1) * wpr;
        const uint32_t *new_dn = cluster.map_shared_rank(s_bits, dn_rank) + (cur ^ 1) * rows * wpr;
        int nup = 0; T rsum = (T)0;
#pragma unroll
        for (int j = 0; j < kIsingRB; j++) {
            const int r = band * kIsingRB + j;
            if (r < rows && active) {
                const int s = keep_sa[j] & 7, a = keep_sa[j] >> 3;
                const int up = r == 0 ? (int)((new_up[x >> 5] >> (x & 31)) & 1u) : bit_at(new_own, wpr, r - 1, x);
                const int dn = r == rows - 1 ? (int)((new_dn[x >> 5] >> (x & 31)) & 1u) : bit_at(new_own, wpr, r + 1, x);
                const int ups = up + dn + bit_at(new_own, wpr, r, xl) + bit_at(new_own, wpr, r, xr);
                const T reward = (T)0.5 * (T)(2 * a - 1) * (T)(2 * ups - 4);
                s_q[((size_t)s * strip + r * L + x) * 2 + a] = keep_q[j] + A.lr * (reward - keep_q[j]);
                nup += a; rsum += reward;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nup += __shfl_xor_sync(0xFFFFFFFFu, nup, o);
            rsum += __shfl_xor_sync(0xFFFFFFFFu, rsum, o);
        }
        if (lane == 0) { atomicAdd(&s_red_i, nup); if (A.reward_sum) atomicAdd(&s_red_r, rsum); }
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&A.n_up[(size_t)k * A.B + b], s_red_i);
            if (A.reward_sum) atomicAdd(&A.reward_sum[(size_t)k * A.B + b], s_red_r);
        }
        __syncthreads();   // s_red_* are rewritten at the top of the next sweep
    }

    // ---- store: Q strip and the final spins ----
    for (int sp = 0; sp < 5; sp++) {
        T *dst = A.Q + (lbase * 5 + (size_t)sp * N + (size_t)row0 * L) * 2;
        const T *src = s_q + (size_t)sp * strip * 2;
        for (int i = tid; i < strip * 2; i += blockDim.x) dst[i] = src[i];
    }
    const uint32_t *fin = s_bits + cur * rows * wpr;
#pragma unroll
    for (int j = 0; j < kIsingRB; j++) {
        const int r = band * kIsingRB + j;
        if (r < rows && active) A.spins[lbase + (size_t)(row0 + r) * L + x] = (int8_t)bit_at(fin, wpr, r, x);
    }
    cluster.sync();       // nobody leaves while a neighbour may still read its shared memory
}

// cluster size for a lattice side, 0 = the resident kernel does not apply
template <typename T>
static int resident_cluster_size(int L) {
    const int wpr = (L + 31) >> 5, LP = wpr * 32;
    for (int C = 1; C <= 16; C *= 2) {
        if (L % C) continue;
        const int rows = L / C;
        if (C > 1 && rows % kIsingRB) continue;
        const int bands = (rows + kIsingRB - 1) / kIsingRB;
        const size_t smem = (size_t)5 * rows * L * 2 * sizeof(T) + (size_t)2 * rows * wpr * 4;
        if (LP * bands <= 1024 && smem <= 200 * 1024) return C;
    }
    return 0;
}

template <typename T>
static void launch_ising_resident(const IsingRunArgs<T> &A0, cudaStream_t st) {
    IsingRunArgs<T> A = A0;
    const int L = A.L, wpr = (L + 31) >> 5, LP = wpr * 32;
    const int C = resident_cluster_size<T>(L);
    if (C == 0) throw Fatal("ising resident kernel: lattice side " + std::to_string(L) + " not supported (use mfi_step)");
    A.rows_per = L / C;
    const int bands = (A.rows_per + kIsingRB - 1) / kIsingRB;
    const size_t smem = (size_t)5 * A.rows_per * L * 2 * sizeof(T) + (size_t)2 * A.rows_per * wpr * 4;
    MF_CUDA(cudaFuncSetAttribute(k_ising_resident<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (C > 8) MF_CUDA(cudaFuncSetAttribute(k_ising_resident<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(A.B * C)); cfg.blockDim = dim3((unsigned)(LP * bands));
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    MF_CUDA(cudaLaunchKernelEx(&cfg, k_ising_resident<T>, A));
}

}  // namespace mfmarl

using namespace mfmarl;

extern "C" int mfi_step(int dtype, int n_lattices, int side, int8_t *d_spins, void *d_q, double temperature,
                        double lr, const void *d_uniforms, const uint8_t *d_update_mask, unsigned seed,
                        unsigned lattice_base, unsigned step, int32_t *d_n_up, void *d_reward_sum, void *d_mse,
                        void *stream) {
    try {
        if (dtype == 0) {
            IsingArgs<float> A{n_lattices, side, d_spins, (float *)d_q, (float)temperature, (float)lr,
                               (const float *)d_uniforms, d_update_mask, seed, lattice_base, step, d_n_up,
                               (float *)d_reward_sum, (float *)d_mse};
            launch_ising(A, (cudaStream_t)stream);
        } else if (dtype == 1) {
            IsingArgs<double> A{n_lattices, side, d_spins, (double *)d_q, temperature, lr,
                                (const double *)d_uniforms, d_update_mask, seed, lattice_base, step, d_n_up,
                                (double *)d_reward_sum, (double *)d_mse};
            launch_ising(A, (cudaStream_t)stream);
        } else throw Fatal("mfi_step: dtype must be 0 (f32) or 1 (f64)");
    } catch (const std::exception &ex) {
        set_last_error(std::string("mfi_step: ") + ex.what());
        return -1;
    }
    return 0;
}

extern "C" int mfi_resident_cluster_size(int dtype, int side) {
    return dtype == 1 ? resident_cluster_size<double>(side) : resident_cluster_size<float>(side);
}

extern "C" int mfi_run(int dtype, int n_lattices, int side, int n_sweeps, int8_t *d_spins, void *d_q,
                       const void *d_temperatures, double lr, const void *d_uniforms, unsigned seed,
                       unsigned lattice_base, unsigned step0, int32_t *d_n_up, void *d_reward_sum, void *stream) {
    try {
        if (n_sweeps < 1 || n_lattices < 1) throw Fatal("mfi_run: need at least one sweep and one lattice");
        if (dtype == 0) {
            IsingRunArgs<float> A{n_lattices, side, n_sweeps, 0, d_spins, (float *)d_q, (const float *)d_temperatures,
                                  (float)lr, (const float *)d_uniforms, seed, lattice_base, step0, d_n_up,
                                  (float *)d_reward_sum};
            launch_ising_resident(A, (cudaStream_t)stream);
        } else if (dtype == 1) {
            IsingRunArgs<double> A{n_lattices, side, n_sweeps, 0, d_spins, (double *)d_q, (const double *)d_temperatures,
                                   lr, (const double *)d_uniforms, seed, lattice_base, step0, d_n_up,
                                   (double *)d_reward_sum};
            launch_ising_resident(A, (cudaStream_t)stream);
        } else throw Fatal("mfi_run: dtype must be 0 (f32) or 1 (f64)");
    } catch (const std::exception &ex) {
        set_last_error(std::string("mfi_run: ") + ex.what());
        return -1;
    }
    return 0;
}
