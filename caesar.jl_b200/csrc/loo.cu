// loo.cu -- per-dimension leave-one-out likelihood cross-validation bandwidth (K4): what upstream
// `manikde!(M, pts)` does when no bandwidth is given (call site R:test/testScatterAlignPose2.jl:187).
//
// A "problem" is one coordinate of one density: N scalars.  The golden-section search state of every problem
// lives on the device; each iteration is two launches for the whole batch (no host round trip):
//   loo_sort_init_kernel : once: sorts the coordinate (shared-memory bitonic sort, FP64), derives the bracket from
//                          the smallest neighbour gap and the range
//   loo_rows_kernel      : for every active problem the row sums sum_{j != i} exp2(-(x_i-x_j)^2 c(h)), one CTA per
//                          (problem, block of 1024 sorted rows), rows in registers, columns broadcast from shared
//                          memory.  On sorted data a column tile that is farther than the flush-to-zero radius from
//                          the whole row block contributes exactly 0 and is skipped (bit-identical result).  The
//                          kernel is symmetric, so a row block only visits the column tiles at or above its own and
//                          hands the column sums down to the later row blocks (loo_finish_kernel adds them, takes logs).
//   golden_step_kernel   : one thread per problem folds the row-block partials in fixed order, turns them into
//                          the objective value and advances that problem's bracket.
// The bracket and the update rule follow Ihler's `ksize(..., 'lcv')` / `golden` as restated in oracle/cb2_oracle.c.
#include "common.cuh"
#include <type_traits>
#include <cooperative_groups.h>

namespace {

constexpr int kThreads = 256;
constexpr int kRI = 4;
#ifndef CB2_LOO_SOFT_EVERY
#define CB2_LOO_SOFT_EVERY 8
#endif
constexpr int kSoftEvery = CB2_LOO_SOFT_EVERY;  // multiple of kRI
constexpr int kCT = 1024;
constexpr int kSortThreads = 1024;
constexpr double kTol = 1e-2;
constexpr int kSmallN = 1024;

struct LooProblem {
  const double* x;  // first scalar of this coordinate
  int stride;       // distance between consecutive points (= d)
  int N;
  int circular;
  int part0;        // slot of this problem in the partial array
  int nrb;
  long long xs_off; // offset of this problem's sorted copy in the xs scratch (also of its row sums)
  int nstride;      // padded N: stride of one row block's column-sum slice
  long long cp_off; // offset of this problem's column-sum slices [nrb][nstride]
  int small;        // N <= kSmallN: solved by the single-CTA kernel
};

struct LooState {
  double x0, x1, x2, x3, f1, f2;
  double h_eval;  // bandwidth to evaluate in the next rows pass
  int phase;      // 0: evaluating x1, 1: evaluating x2, 2: iterating (evaluating the new interior point), 3: done
  int slot;       // which of (x1,x2) h_eval belongs to while iterating
  double result;
};

__device__ __forceinline__ void loo_bracket(LooState& s, double min_gap, double range);

// one CTA per problem: sorted copy + golden-section bracket
__global__ void __launch_bounds__(kSortThreads)
loo_sort_init_kernel(const LooProblem* __restrict__ probs, LooState* __restrict__ st, double* __restrict__ xs_all) {
  extern __shared__ __align__(16) double skey[];
  __shared__ double sred[kSortThreads / 32];
  const LooProblem pr = probs[blockIdx.x];
  if (pr.small) return;  // handled entirely by loo_small_kernel
  const int N = pr.N, tid = threadIdx.x, nthr = blockDim.x;
  int P2 = 1;
  while (P2 < N) P2 <<= 1;
  for (int i = tid; i < P2; i += nthr) skey[i] = i < N ? pr.x[(size_t)i * pr.stride] : 1.7976931348623157e308;
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P2; i += nthr) {
        int ixj = i ^ j;
        if (ixj > i) {
          double a = skey[i], b = skey[ixj];
          bool up = (i & k) == 0;
          if ((b < a) == up) { skey[i] = b; skey[ixj] = a; }
        }
      }
      __syncthreads();
    }
  double* xs = xs_all + pr.xs_off;
  double mn = 1e300;
  for (int i = tid; i < N; i += nthr) {
    xs[i] = skey[i];
    if (i + 1 < N) mn = fmin(mn, skey[i + 1] - skey[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((tid & 31) == 0) sred[tid >> 5] = mn;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < nthr / 32; ++w) mn = fmin(mn, sred[w]);
    LooState s;
    loo_bracket(s, mn, skey[N - 1] - skey[0]);
    st[blockIdx.x] = s;
  }
}

template <typename T> struct ZeroRadius;  // |scaled difference| beyond which exp2(-d^2) is exactly 0 in T
template <> struct ZeroRadius<float> { static constexpr double value = 11.4; };   // d^2 > 129: below FP32 min normal (ftz)
template <> struct ZeroRadius<double> { static constexpr double value = 33.0; };  // d^2 > 1089: below FP64 denormals

// sum over the 32 lanes of 32 per-lane values in 31 shuffles: afterwards lane L holds the total of v[L]
template <typename T> __device__ __forceinline__ T lane_transpose_sum(T (&v)[32], int lane) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int off = 16 >> s;
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      T send = hi ? v[k] : v[k + off];
      T keep = hi ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// All pairs among up to 1024 sorted points held in shared memory (`sxs`, padded with a far sentinel up to a multiple of 256;
// thread t owns rows t + 256 r, xi / acc in registers).  The points form up to four blocks of 256.  Pairs inside a block are
// evaluated in full with the self term excluded (it sits in the 32-column group at the warp's own base); pairs between blocks
// a < b are evaluated ONCE: k goes to the row of block a (register) and to the column of block b (lane transpose + per-warp
// shared-memory slots `colacc` [warps][256], summed in fixed order by the thread that owns that row): 0.62 of the
// exponentials at n = 1000.  WRAP = circular coordinate that may wrap (four more instructions per pair).
template <typename T, bool WRAP>
__device__ __forceinline__ void block_pairs_symmetric(const T* __restrict__ sxs, int n, const T (&xi)[kRI], T (&acc)[kRI], T period,
                                                      T* __restrict__ colacc, int tid) {
  const int tid_base = tid & ~31, lane = tid & 31, warp = tid >> 5;
  const int nb = (n + kThreads - 1) / kThreads;
#pragma unroll
  for (int r = 0; r < kRI; ++r) {
    if (r >= nb) break;
    const int c0 = r * kThreads, c1 = min(n, c0 + kThreads);
    for (int j0 = c0; j0 < c1; j0 += 32) {
      const int je = min(j0 + 32, c1);
      if (j0 - c0 == tid_base) {
        for (int j = j0; j < je; ++j) {
          T df = xi[r] - sxs[j];
          if (WRAP) { T ad = fabs(df); df = fmin(ad, period - ad); }
          T kv = fast_exp2(-(df * df));
          acc[r] += (j == tid + c0) ? T(0) : kv;
        }
      } else {
#pragma unroll 8
        for (int j = j0; j < je; ++j) {
          T df = xi[r] - sxs[j];
          if (WRAP) { T ad = fabs(df); df = fmin(ad, period - ad); }
          acc[r] += fast_exp2(-(df * df));
        }
      }
    }
  }
#pragma unroll
  for (int bb = 1; bb < kRI; ++bb) {
    if (bb >= nb) break;
    const int c0 = bb * kThreads, ncol = min(n, c0 + kThreads) - c0;
#pragma unroll
    for (int aa = 0; aa < bb; ++aa) {
      __syncthreads();  // colacc is free again
      for (int j0 = 0; j0 < kThreads; j0 += 32) {
        T v[32];
        if constexpr (!WRAP && sizeof(T) == 4) {
          // two columns per packed operation: (q_c, q_c+1) comes out of shared memory as one 64-bit broadcast load
          const f32x2 nx = pack2(-(float)xi[aa], -(float)xi[aa]);
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            f32x2 qq;
            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(qq) : "r"((uint32_t)__cvta_generic_to_shared(&sxs[c0 + j0 + c])));
            const f32x2 d2 = add2(qq, nx);
            float sa, sb;
            unpack2(mul2(d2, d2), sa, sb);
            v[c] = (T)fast_exp2(-sa);      // columns past n hold the far sentinel: exp2 -> 0
            v[c + 1] = (T)fast_exp2(-sb);
            acc[aa] += v[c];
            acc[aa] += v[c + 1];
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            T df = xi[aa] - sxs[c0 + j0 + c];  // columns past n hold the far sentinel: exp2 -> 0 without the wrap
            if (WRAP) { T ad = fabs(df); df = fmin(ad, period - ad); }
            T kv = fast_exp2(-(df * df));
            if (WRAP && j0 + c >= ncol) kv = T(0);
            acc[aa] += kv;
            v[c] = kv;
          }
        }
        colacc[warp * kThreads + j0 + lane] = lane_transpose_sum(v, lane);
      }
      __syncthreads();
      T t = T(0);
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) t += colacc[w * kThreads + tid];
      acc[bb] += t;
    }
  }
}

// k(x_i, x_j) is symmetric: the CTA of row block I only visits column tiles J >= I.  The diagonal tile is evaluated in
// full; for a tile above the diagonal every kernel value is added to its row (registers) AND to its column (lane
// transpose + per-warp shared-memory slots, summed in fixed order), and the column sums are handed to the rows of
// block J through `colpart`.  Half the MUFU work of the plain row pass; deterministic (no atomics).
template <typename T>
__global__ void __launch_bounds__(kThreads)
loo_rows_kernel(const LooProblem* __restrict__ probs, const LooState* __restrict__ st, const int2* __restrict__ cta_map,
                const double* __restrict__ xs_all, T* __restrict__ rowsum_all, T* __restrict__ colpart_all) {
  const int2 cm = cta_map[blockIdx.x];  // (problem, row block)
  const LooState s = st[cm.x];
  if (s.phase == 3) return;
  const LooProblem pr = probs[cm.x];
  const double* __restrict__ xs = xs_all + pr.xs_off;
  extern __shared__ __align__(16) unsigned char loo_smem[];
  T* sx = reinterpret_cast<T*>(loo_smem);           // [kCT]
  T* colacc = sx + kCT;                              // [kThreads/32][kCT]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, N = pr.N;
  // exponent in base 2: (x_i - x_j)^2 * 0.5*log2e/h^2 -> pre-scale the coordinates by sc
  const double sc = sqrt(0.5 * CB2_LOG2E) / s.h_eval;
  const bool circ = pr.circular != 0;
  // common shift keeps FP32 differences accurate; circular coordinates are wrapped to [-pi, pi] instead, so that
  // |x_i - x_j| <= 2 pi and the wrapped distance is min(|d|, 2 pi - |d|)
  const double x_mid = circ ? 0.0 : xs[N >> 1];
  const T period = (T)(CB2_TWO_PI * sc);
  T xi[kRI], acc[kRI];
  int rowi[kRI];
  const int RB = kThreads * kRI;
  const int rbase = cm.y * RB;
  const int rend = min(rbase + RB, N);
#pragma unroll
  for (int r = 0; r < kRI; ++r) {
    int row = rbase + r * kThreads + tid;
    rowi[r] = row;
    double v = xs[row < N ? row : rbase];
    xi[r] = (T)(((circ ? wrap_pi(v) : v) - x_mid) * sc);
    acc[r] = T(0);
  }
  const double row_hi = xs[rend - 1];
  const double zero_r = ZeroRadius<T>::value / sc;
  T* colpart = colpart_all + pr.cp_off + (size_t)cm.y * pr.nstride;  // this row block's slice: [nstride]
  // the pair loops are compiled once per topology: the wrap of circular coordinates costs four instructions per pair, which the
  // Euclidean coordinates must not pay (the kernel sits at the issue limit next to the XU limit); uniform branch per CTA
  auto sweep = [&](auto topo) {
    constexpr bool CIRC = decltype(topo)::value;
    for (int cs = rbase; cs < N; cs += kCT) {
      const int cend = min(cs + kCT, N);
      const bool diag = cs == rbase;
      if (!CIRC && !diag && xs[cs] - row_hi > zero_r) {
        // sorted data: a tile entirely beyond the flush-to-zero radius contributes exactly 0 (and so do all later ones)
        for (int j = tid; j < cend - cs; j += kThreads) colpart[cs + j] = T(0);
        continue;
      }
      __syncthreads();
      for (int j = tid; j < kCT; j += kThreads) {
        int col = cs + j;
        if (col < N) { double v = xs[col]; sx[j] = (T)(((CIRC ? wrap_pi(v) : v) - x_mid) * sc); }
        else sx[j] = T(1e18);
      }
      __syncthreads();
      const int jn = cend - cs;
      // sorted data: when every coordinate of the tile's rows and columns lies in [-pi, pi] within pi of each other, the wrapped
      // distance is the plain one and the tile takes the Euclidean arithmetic (same values, two instructions per pair fewer);
      // beliefs that do not straddle the +-pi cut never wrap
      bool wrap = false;
      if (CIRC) {
        const double lo = xs[rbase], hi = xs[cend - 1];
        wrap = !(lo >= -CB2_PI && hi <= CB2_PI && hi - lo <= CB2_PI);
      }
      auto pairs = [&](auto wrap_c) {
        constexpr bool WRAP = decltype(wrap_c)::value;
        if (diag) {  // all pairs inside the block, self terms excluded, pairs between its 256-row sub-blocks evaluated once
          block_pairs_symmetric<T, WRAP>(sx, jn, xi, acc, period, colacc, tid);
        } else {  // tile above the diagonal: the row block is complete (1024 valid rows); columns may be ragged
          // only the last 32-column group of a ragged tile needs the padding mask (circular case: the wrap would fold the
          // sentinel back into range); every other group runs without it
          auto colgroup = [&](int j0, auto ragged) {
            constexpr bool RAGGED = decltype(ragged)::value;
            (void)RAGGED;
            T v[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const T q = sx[j0 + c];
              T kv[kRI];
              if constexpr (!WRAP && sizeof(T) == 4 && kRI == 4) {
                // the four rows as two packed pairs: 2 FADD2 + 2 FMUL2 instead of 4 FADD + 4 FMUL per column
                const f32x2 qq = pack2((float)q, (float)q);
                const f32x2 d01 = add2(qq, pack2(-(float)xi[0], -(float)xi[1])), d23 = add2(qq, pack2(-(float)xi[2], -(float)xi[3]));
                float s[4];
                unpack2(mul2(d01, d01), s[0], s[1]);
                unpack2(mul2(d23, d23), s[2], s[3]);
#pragma unroll
                for (int r = 0; r < kRI; ++r) {
                  // XU relief: one evaluation in kSoftEvery runs on the FMA/ALU pipes (soft_exp2) where the issue rate leaves room
                  kv[r] = (T)((r == kRI - 1 && (c % (kSoftEvery / kRI)) == 0) ? soft_exp2(-s[r]) : fast_exp2(-s[r]));
                  acc[r] += kv[r];
                }
              } else {
#pragma unroll
                for (int r = 0; r < kRI; ++r) {
                  T df = xi[r] - q;
                  if (WRAP) { T ad = fabs(df); df = fmin(ad, period - ad); }
                  kv[r] = fast_exp2(-(df * df));
                  if (WRAP && RAGGED && j0 + c >= jn) kv[r] = T(0);
                  acc[r] += kv[r];
                }
              }
              v[c] = (kv[0] + kv[1]) + (kv[2] + kv[3]);
            }
            colacc[warp * kCT + j0 + lane] = lane_transpose_sum(v, lane);
          };
          for (int j0 = 0; j0 < kCT; j0 += 32) {
            if (WRAP && j0 + 32 > jn) colgroup(j0, std::true_type{}); else colgroup(j0, std::false_type{});
          }
          __syncthreads();
          for (int j = tid; j < jn; j += kThreads) {
            T t = T(0);
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) t += colacc[w * kCT + j];
            colpart[cs + j] = t;
          }
        }
      };
      if (CIRC && wrap) pairs(std::true_type{}); else pairs(std::false_type{});
    }
  };
  if (circ) sweep(std::true_type{}); else sweep(std::false_type{});
  T* rowsum = rowsum_all + pr.xs_off;
#pragma unroll
  for (int r = 0; r < kRI; ++r)
    if (rowi[r] < N) rowsum[rowi[r]] = acc[r];
}

// row totals = own row pass + the column sums handed down by the row blocks above; then sum_i log(.)
template <typename T>
__global__ void __launch_bounds__(kThreads)
loo_finish_kernel(const LooProblem* __restrict__ probs, const LooState* __restrict__ st, const T* __restrict__ rowsum_all,
                  const T* __restrict__ colpart_all, double* __restrict__ partial) {
  const int p = blockIdx.x;
  if (st[p].phase == 3) return;
  const LooProblem pr = probs[p];
  __shared__ double sred[kThreads / 32];
  const T* rowsum = rowsum_all + pr.xs_off;
  const T* colpart = colpart_all + pr.cp_off;
  const int RB = kThreads * kRI;
  double v = 0;
  for (int r = threadIdx.x; r < pr.N; r += kThreads) {
    double a = (double)rowsum[r];
    const int I = r / RB;
    for (int b = 0; b < I; ++b) a += (double)colpart[(size_t)b * pr.nstride + r];
    v += log(a > 2.2250738585072014e-308 ? a : 2.2250738585072014e-308);
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += sred[w];
    partial[p] = t;
  }
}

// one golden-section step (Numerical-Recipes form used by Ihler's `golden`): f is the objective at s.h_eval
__device__ __forceinline__ void golden_advance(LooState& s, double f) {
  const double C = (3.0 - sqrt(5.0)) / 2.0, R = 1.0 - C;
  if (s.phase == 0) { s.f1 = f; s.phase = 1; s.h_eval = s.x2; return; }
  if (s.phase == 1) { s.f2 = f; s.phase = 2; }
  else if (s.slot == 2) s.f2 = f;
  else s.f1 = f;
  if (fabs(s.x3 - s.x0) > kTol * (fabs(s.x1) + fabs(s.x2))) {
    if (s.f2 < s.f1) { s.x0 = s.x1; s.x1 = s.x2; s.x2 = R * s.x1 + C * s.x3; s.f1 = s.f2; s.h_eval = s.x2; s.slot = 2; }
    else { s.x3 = s.x2; s.x2 = s.x1; s.x1 = R * s.x2 + C * s.x0; s.f2 = s.f1; s.h_eval = s.x1; s.slot = 1; }
  } else {
    s.result = s.f1 < s.f2 ? s.x1 : s.x2;
    s.phase = 3;
  }
}

// nll(h) = -(1/N) sum_i [ log(sum_{j!=i} k_ij) - 0.5 log(2 pi h^2) - log(N-1) ]
__device__ __forceinline__ double loo_objective(double sum_log_rows, int N, double h) {
  return -(sum_log_rows / N - 0.5 * log(CB2_TWO_PI * h * h) - log((double)(N - 1)));
}

__device__ __forceinline__ void loo_bracket(LooState& s, double min_gap, double range) {
  double m = min_gap < 1e-6 ? 1e-6 : min_gap;
  if (!(range >= 1e-6)) {
    s.phase = 3; s.result = 1e-6; s.h_eval = 1.0;
    s.x0 = s.x1 = s.x2 = s.x3 = s.f1 = s.f2 = 0; s.slot = 0;
    return;
  }
  const double C = (3.0 - sqrt(5.0)) / 2.0;
  double ax = 2.0 * m / (1.0 + sqrt(5.0)), bx = m, cx = 2.0 * range;
  s.x0 = ax; s.x3 = cx;
  if (fabs(cx - bx) > fabs(bx - ax)) { s.x1 = bx; s.x2 = bx + C * (cx - bx); }
  else { s.x2 = bx; s.x1 = bx - C * (bx - ax); }
  s.f1 = s.f2 = 0; s.phase = 0; s.slot = 1; s.h_eval = s.x1; s.result = 0;
}

__global__ void golden_step_kernel(const LooProblem* __restrict__ probs, int P, const double* __restrict__ partial,
                                   LooState* __restrict__ st, int* __restrict__ n_active) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  LooState s = st[p];
  if (s.phase == 3) return;
  golden_advance(s, loo_objective(partial[p], probs[p].N, s.h_eval));
  st[p] = s;
  if (s.phase != 3) atomicAdd(n_active, 1);
}

// Small problems (N <= kSmallN, the usual solver particle counts 100..1000): the whole search -- sort, bracket, every
// golden-section step -- runs inside ONE CTA per problem, no per-iteration launches (latency: microseconds, not ~0.4 ms).
template <typename T>
__global__ void __launch_bounds__(kThreads)
loo_small_kernel(const LooProblem* __restrict__ probs, const int* __restrict__ ids, LooState* __restrict__ st) {
  const int p = ids[blockIdx.x];
  const LooProblem pr = probs[p];
  const int N = pr.N, tid = threadIdx.x;
  __shared__ double sk[kSmallN];
  __shared__ T sxs[kSmallN];
  __shared__ T colacc[(kThreads / 32) * kThreads];  // per-warp column sums of one block pair
  __shared__ double sred[kThreads / 32];
  __shared__ LooState ss;
  int P2 = 1;
  while (P2 < N) P2 <<= 1;
  for (int i = tid; i < P2; i += kThreads) sk[i] = i < N ? pr.x[(size_t)i * pr.stride] : 1.7976931348623157e308;
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P2; i += kThreads) {
        int ixj = i ^ j;
        if (ixj > i) {
          double a = sk[i], b = sk[ixj];
          if ((b < a) == ((i & k) == 0)) { sk[i] = b; sk[ixj] = a; }
        }
      }
      __syncthreads();
    }
  double mn = 1e300;
  for (int i = tid; i + 1 < N; i += kThreads) mn = fmin(mn, sk[i + 1] - sk[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((tid & 31) == 0) sred[tid >> 5] = mn;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kThreads / 32; ++w) mn = fmin(mn, sred[w]);
    LooState s0;
    loo_bracket(s0, mn, sk[N - 1] - sk[0]);
    ss = s0;
  }
  const bool circ = pr.circular != 0;
  const double x_mid = circ ? 0.0 : sk[N >> 1];
  // sorted: a circular problem whose points lie in [-pi, pi] within pi of each other never needs the wrap
  const bool wraps = circ && !(sk[0] >= -CB2_PI && sk[N - 1] <= CB2_PI && sk[N - 1] - sk[0] <= CB2_PI);
  for (int it = 0; it < 96; ++it) {
    __syncthreads();
    if (ss.phase == 3) break;
    const double h = ss.h_eval;
    const double sc = sqrt(0.5 * CB2_LOG2E) / h;
    const T period = (T)(CB2_TWO_PI * sc);
    for (int j = tid; j < kSmallN; j += kThreads) sxs[j] = j < N ? (T)(((circ ? wrap_pi(sk[j]) : sk[j]) - x_mid) * sc) : T(1e18);
    __syncthreads();
    T xi[kRI], acc[kRI];
#pragma unroll
    for (int r = 0; r < kRI; ++r) { int row = tid + r * kThreads; xi[r] = sxs[row < N ? row : 0]; acc[r] = T(0); }
    if (wraps) block_pairs_symmetric<T, true>(sxs, N, xi, acc, period, colacc, tid);
    else block_pairs_symmetric<T, false>(sxs, N, xi, acc, period, colacc, tid);
    double v = 0;
#pragma unroll
    for (int r = 0; r < kRI; ++r)
      if (tid + r * kThreads < N) {
        double a = (double)acc[r];
        v += log(a > 2.2250738585072014e-308 ? a : 2.2250738585072014e-308);
      }
    v = warp_sum(v);
    if ((tid & 31) == 0) sred[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < kThreads / 32; ++w) t += sred[w];
      LooState s1 = ss;
      golden_advance(s1, loo_objective(t, N, h));
      ss = s1;
    }
  }
  __syncthreads();
  if (tid == 0) {
    LooState s1 = ss;
    if (s1.phase != 3) { s1.result = s1.f1 < s1.f2 ? s1.x1 : s1.x2; s1.phase = 3; }  // iteration cap: best point so far
    st[p] = s1;
  }
}

// Latency variant of the small-problem search for SMALL BATCHES (a Gibbs sweep of a ten-variable graph asks for a few dozen
// bandwidths at a time: one CTA per problem would leave most of the 148 SMs idle and every search would take the time of N^2 / 2
// exponentials on one SM).  A thread-block CLUSTER of kLooCluster CTAs works on one problem: every CTA holds all N points in
// shared memory and owns N / kLooCluster rows (one row per thread, all N columns, no symmetry hand-down); per golden-section
// step the CTAs exchange their partial sums of log row sums through distributed shared memory (each CTA writes its value into
// every peer's slot, one cluster barrier per step, slots double-buffered by step parity) and advance identical copies of the
// search state.  Each CTA sorts its own copy (the bracket needs the smallest neighbour gap).
constexpr int kLooCluster = 8;
constexpr int kLooClusterThreads = kSmallN / kLooCluster;  // row slots per CTA: one row per slot
constexpr int kLooClusterCta = 2 * kLooClusterThreads;     // threads per CTA: every row is split into two column segments
template <typename T>
__global__ void __cluster_dims__(kLooCluster, 1, 1) __launch_bounds__(kLooClusterCta)
loo_cluster_kernel(const LooProblem* __restrict__ probs, const int* __restrict__ ids, LooState* __restrict__ st) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int p = ids[blockIdx.x / kLooCluster];
  const LooProblem pr = probs[p];
  const int N = pr.N, tid = threadIdx.x;
  __shared__ double sk[kSmallN];
  __shared__ __align__(16) T sxs[kSmallN + 4];
  __shared__ T part[kLooClusterCta];
  __shared__ double sred[kLooClusterCta / 32];
  __shared__ double slots[2][kLooCluster];
  __shared__ LooState ss;
  int P2 = 1;
  while (P2 < N) P2 <<= 1;
  for (int i = tid; i < P2; i += kLooClusterCta) sk[i] = i < N ? pr.x[(size_t)i * pr.stride] : 1.7976931348623157e308;
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P2; i += kLooClusterCta) {
        int ixj = i ^ j;
        if (ixj > i) {
          double a = sk[i], b = sk[ixj];
          if ((b < a) == ((i & k) == 0)) { sk[i] = b; sk[ixj] = a; }
        }
      }
      __syncthreads();
    }
  double mn = 1e300;
  for (int i = tid; i + 1 < N; i += kLooClusterCta) mn = fmin(mn, sk[i + 1] - sk[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((tid & 31) == 0) sred[tid >> 5] = mn;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kLooClusterCta / 32; ++w) mn = fmin(mn, sred[w]);
    LooState s0;
    loo_bracket(s0, mn, sk[N - 1] - sk[0]);
    ss = s0;
  }
  const bool circ = pr.circular != 0;
  const double x_mid = circ ? 0.0 : sk[N >> 1];
  const bool wraps = circ && !(sk[0] >= -CB2_PI && sk[N - 1] <= CB2_PI && sk[N - 1] - sk[0] <= CB2_PI);
  const int rows = (N + kLooCluster - 1) / kLooCluster;   // rows of this problem per CTA (<= kLooClusterThreads)
  // (row slot, column segment): with one thread per row a scheduler holds ONE warp and every dependent step of a pair
  // (load, subtract, square, exponential, add) is exposed -- 24.7 k clk per golden-section step at N = 1000 against 8 k clk of
  // exponentials; two segments per row give every scheduler a second warp, and the FP32 loop reads four columns per LDS.128 and
  // squares them as packed pairs
  const int rl = tid & (kLooClusterThreads - 1), seg = tid / kLooClusterThreads;
  const int row = rank * rows + rl;
  const bool mine = rl < rows && row < N;
  const int half = (((N + 1) >> 1) + 3) & ~3;              // first column of the second segment: a multiple of four
  const int c0 = seg ? min(N, half) : 0, c1 = seg ? N : min(N, half);
  const int c1p = (c1 + 3) & ~3;                           // the last block of four may run into the padding behind N
  const int npad = (N + 3) & ~3;
#ifdef CB2_LOO_PROF
  long long cp[5] = {0, 0, 0, 0, 0}, cl = clock64();
  int csteps = 0;
#define LOOC_T(k) do { long long t_ = clock64(); cp[k] += t_ - cl; cl = t_; } while (0)
#else
#define LOOC_T(k) do { } while (0)
#endif
  for (int it = 0; it < 96; ++it) {
    __syncthreads();
    if (ss.phase == 3) break;  // identical in every CTA of the cluster: all leave together
    LOOC_T(4);
    const double h = ss.h_eval;
    const double sc = sqrt(0.5 * CB2_LOG2E) / h;
    const T period = (T)(CB2_TWO_PI * sc);
    for (int j = tid; j < npad; j += kLooClusterCta) sxs[j] = j < N ? (T)(((circ ? wrap_pi(sk[j]) : sk[j]) - x_mid) * sc) : T(1e18);
    __syncthreads();
    LOOC_T(0);
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
    if (mine && c1 > c0) {
      const T xi = sxs[row];
      if (wraps) {
        for (int j = c0; j < c1; ++j) {
          T ad = fabs(xi - sxs[j]);
          T df = fmin(ad, period - ad);
          a0 += (j == row) ? T(0) : fast_exp2(-(df * df));
        }
      } else {
        if constexpr (sizeof(T) == 4) {
          const f32x2 nx = pack2(-(float)xi, -(float)xi);
#pragma unroll 4
          for (int j = c0; j < c1p; j += 4) {   // the self term (exp2(0) = 1) is summed and taken out below
            f32x2 q01, q23;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(q01), "=l"(q23) : "r"((uint32_t)__cvta_generic_to_shared(&sxs[j])));
            const f32x2 d01 = add2(q01, nx), d23 = add2(q23, nx);
            float s0, s1, s2, s3;
            unpack2(mul2(d01, d01), s0, s1);
            unpack2(mul2(d23, d23), s2, s3);
            a0 += fast_exp2(-s0); a1 += fast_exp2(-s1); a2 += fast_exp2(-s2); a3 += fast_exp2(-s3);
          }
        } else {
#pragma unroll 2
          for (int j = c0; j < c1p; j += 4) {
            const T d0 = xi - sxs[j], d1 = xi - sxs[j + 1], d2 = xi - sxs[j + 2], d3 = xi - sxs[j + 3];
            a0 += fast_exp2(-(d0 * d0)); a1 += fast_exp2(-(d1 * d1)); a2 += fast_exp2(-(d2 * d2)); a3 += fast_exp2(-(d3 * d3));
          }
        }
        if (row >= c0 && row < c1) a0 -= T(1);
      }
    }
    part[tid] = (a0 + a1) + (a2 + a3);
    LOOC_T(1);
    __syncthreads();
    if (seg == 0) {
      double v = 0;
      if (mine) {
        const double a = (double)(part[rl] + part[kLooClusterThreads + rl]);
        v = log(a > 2.2250738585072014e-308 ? a : 2.2250738585072014e-308);
      }
      v = warp_sum(v);
      if ((tid & 31) == 0) sred[tid >> 5] = v;
    }
    __syncthreads();
    LOOC_T(2);
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < kLooClusterThreads / 32; ++w) t += sred[w];
      for (int r = 0; r < kLooCluster; ++r) cluster.map_shared_rank(&slots[it & 1][0], r)[rank] = t;  // my partial into every peer
    }
    cluster.sync();
    LOOC_T(3);
    if (tid == 0) {
      double t = 0;
      for (int r = 0; r < kLooCluster; ++r) t += slots[it & 1][r];  // fixed order: every CTA computes the same total
      LooState s1 = ss;
      golden_advance(s1, loo_objective(t, N, h));
      ss = s1;
    }
#ifdef CB2_LOO_PROF
    ++csteps;
#endif
  }
#ifdef CB2_LOO_PROF
  if (blockIdx.x == 0 && (tid == 0 || tid == 200))
    printf("loo cluster prof tid %d N %d steps %d clk/step: fill %lld pairs %lld owner %lld exchange %lld advance+barrier %lld\n", tid, N, csteps,
           cp[0] / max(csteps, 1), cp[1] / max(csteps, 1), cp[2] / max(csteps, 1), cp[3] / max(csteps, 1), cp[4] / max(csteps, 1));
#endif
  __syncthreads();
  if (tid == 0 && rank == 0) {
    LooState s1 = ss;
    if (s1.phase != 3) { s1.result = s1.f1 < s1.f2 ? s1.x1 : s1.x2; s1.phase = 3; }  // iteration cap: best point so far
    st[p] = s1;
  }
  cluster.sync();  // no CTA leaves while a peer may still write into its shared memory
}

// Latency variant for the smallest problems (N <= kTinyN: the N = 100 particles IIF solves with by default).  The single-CTA kernel
// above spends ~2.6 us per golden-section step on 100 points -- a thread per row leaves 60 % of the CTA idle, every step rebuilds
// a 1024-entry scaled copy and passes the search state through shared memory behind four barriers.  Here the CTA's 256 threads
// are laid out as (row, column segment): N = 100 gives 128 row slots x 2 segments, N <= 64 four segments, N <= 32 eight; the
// points are stored once (centred, unscaled: the scale of the step is applied to the difference), the partial row sums meet in
// shared memory, and EVERY thread advances an identical private copy of the search state from the per-warp sums, so a step
// costs two barriers.  The bandwidths land in the output array directly (no collect launch).  Same golden-section path as the
// other kernels; sums are taken in a different order (FP32: 1e-6 relative on the objective).
constexpr int kTinyN = 128;
template <typename T>
__global__ void __launch_bounds__(kThreads)
loo_tiny_kernel(const LooProblem* __restrict__ probs, const int* __restrict__ ids, LooState* __restrict__ st, double* __restrict__ out) {
  const int p = ids[blockIdx.x];
  const LooProblem pr = probs[p];
  const int N = pr.N, tid = threadIdx.x;
  __shared__ double sk[kTinyN];
  __shared__ __align__(16) T sx[kTinyN + 4];
  __shared__ T part[kThreads];
  __shared__ double sred[kThreads / 32];
  __shared__ double slogh;
  int P2 = 2;
  while (P2 < N) P2 <<= 1;
  if (tid < P2) sk[tid] = tid < N ? pr.x[(size_t)tid * pr.stride] : 1.7976931348623157e308;
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int ixj = tid ^ j;
      if (tid < P2 && ixj > tid) {
        const double a = sk[tid], b = sk[ixj];
        if ((b < a) == ((tid & k) == 0)) { sk[tid] = b; sk[ixj] = a; }
      }
      __syncthreads();
    }
  double mn = 1e300;
  if (tid + 1 < N) mn = sk[tid + 1] - sk[tid];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((tid & 31) == 0) sred[tid >> 5] = mn;
  const bool circ = pr.circular != 0;
  const double x_mid = circ ? 0.0 : sk[N >> 1];
  const bool wraps = circ && !(sk[0] >= -CB2_PI && sk[N - 1] <= CB2_PI && sk[N - 1] - sk[0] <= CB2_PI);
  // centred, unscaled copy; the padding up to a multiple of four sits far away (its kernel values are exactly zero)
  if (tid < kTinyN + 4) sx[tid] = tid < N ? (T)((circ ? wrap_pi(sk[tid]) : sk[tid]) - x_mid) : T(1e18);
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) mn = fmin(mn, sred[w]);
  LooState s;
  loo_bracket(s, mn, sk[N - 1] - sk[0]);
  const int NR = (N + 31) & ~31;   // row slots: whole warps
  const int S = kThreads / NR;     // column segments per row
  const int row = tid % NR, seg = tid / NR;
  const int cper = ((N + S - 1) / S + 3) & ~3;   // columns per segment, a multiple of four (16-byte loads)
  const int c0 = min(N, seg * cper), c1 = seg < S ? min(N, c0 + cper) : c0;
  const int c1p = (c1 + 3) & ~3;   // the last segment runs into the padding
  const bool live = row < N && c1 > c0;
  const T xi = sx[row < N ? row : 0];
  const T self = (row >= c0 && row < c1) ? T(1) : T(0);   // exp2(0) of the row's own column, summed below and taken out again
  const double invN = 1.0 / N;
  __syncthreads();  // the sred values of the bracket have been read by everyone
#ifdef CB2_LOO_PROF
  long long tp[4] = {0, 0, 0, 0}, tl = clock64();
  int nsteps = 0;
#define LOO_T(k) do { long long t_ = clock64(); tp[k] += t_ - tl; tl = t_; } while (0)
#else
#define LOO_T(k) do { } while (0)
#endif
  for (int it = 0; it < 96 && s.phase != 3; ++it) {  // s is identical in every thread
    const double h = s.h_eval;
    T sc;
    if constexpr (sizeof(T) == 4) sc = 0.84932180028801904f * __frcp_rn((float)h);  // sqrt(0.5 log2 e) / h
    else sc = (T)(sqrt(0.5 * CB2_LOG2E) / h);
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
    if (live) {
      if (wraps) {
        const T period = (T)CB2_TWO_PI;
        for (int j = c0; j < c1; ++j) {
          const T ad = fabs(xi - sx[j]);
          const T df = fmin(ad, period - ad) * sc;
          a0 += fast_exp2(-(df * df));
        }
      } else {
#pragma unroll 4
        for (int j = c0; j < c1p; j += 4) {
          T q[4];
          if constexpr (sizeof(T) == 4) {
            const float4 v = *reinterpret_cast<const float4*>(&sx[j]);
            q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
          } else {
            const double2 v0 = *reinterpret_cast<const double2*>(&sx[j]), v1 = *reinterpret_cast<const double2*>(&sx[j + 2]);
            q[0] = v0.x; q[1] = v0.y; q[2] = v1.x; q[3] = v1.y;
          }
          const T d0 = (xi - q[0]) * sc, d1 = (xi - q[1]) * sc, d2 = (xi - q[2]) * sc, d3 = (xi - q[3]) * sc;
          a0 += fast_exp2(-(d0 * d0)); a1 += fast_exp2(-(d1 * d1)); a2 += fast_exp2(-(d2 * d2)); a3 += fast_exp2(-(d3 * d3));
        }
      }
      a0 -= self;
    }
    part[tid] = (a0 + a1) + (a2 + a3);
    LOO_T(0);
    __syncthreads();
    LOO_T(1);
    if (seg == 0) {  // the first NR / 32 warps own the rows: log of every row sum, summed per warp
      double v = 0;
      if (row < N) {
        T tot = part[row];
        for (int q = 1; q < S; ++q) tot += part[q * NR + row];
        // (an empty row -- every neighbour beyond the flush-to-zero radius -- weighs log(DBL_MIN) as in the other kernels)
        if constexpr (sizeof(T) == 4) v = tot > 0.f ? (double)(__log2f(tot) * 0.693147180559945309f) : -708.3964185322641;
        else v = log(tot > 2.2250738585072014e-308 ? tot : 2.2250738585072014e-308);
      }
      v = warp_sum(v);
      if ((tid & 31) == 0) sred[tid >> 5] = v;
    } else if (tid == kThreads - 1) {  // meanwhile a thread of a segment warp takes the double-precision logarithm of the step
      slogh = log(h);
    }
    LOO_T(2);
    __syncthreads();
    LOO_T(1);
    double t = 0;
    for (int w = 0; w < NR / 32; ++w) t += sred[w];
    // the objective up to its constants (0.5 log 2 pi + log(N - 1) shift every value alike): nll(h) ~ log h - mean_i log rowsum_i
    golden_advance(s, slogh - t * invN);
    LOO_T(3);
#ifdef CB2_LOO_PROF
    ++nsteps;
#endif
  }
#ifdef CB2_LOO_PROF
  if (blockIdx.x == 0 && (tid == 0 || tid == 255))
    printf("loo tiny prof tid %d N %d steps %d clk/step: pairs %lld barriers %lld owner %lld advance %lld\n", tid, N, nsteps,
           tp[0] / max(nsteps, 1), tp[1] / max(nsteps, 1), tp[2] / max(nsteps, 1), tp[3] / max(nsteps, 1));
#endif
  if (tid == 0) {
    if (s.phase != 3) { s.result = s.f1 < s.f2 ? s.x1 : s.x2; s.phase = 3; }  // iteration cap: best point so far
    st[p] = s;
    if (out) out[p] = s.result;
  }
}

__global__ void loo_collect_kernel(const LooState* __restrict__ st, int P, double* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const LooState s = st[p];
  out[p] = s.phase == 3 ? s.result : (s.f1 < s.f2 ? s.x1 : s.x2);  // iteration cap hit: best point so far
}

}  // namespace

// Device-resident core: B densities (device Float64 d x N[b]), sigma_out_dev is B x d (device).  Enqueues the
// whole search on the context stream; checks a device counter every few iterations to stop early.
int cb2_bw_loo_dev_core(cb2_ctx* ctx, const double* const* mu_dev, const int32_t* N, int B, int d,
                        const uint8_t* circular, double* sigma_out_dev, void* scratch, size_t scratch_bytes,
                        size_t* scratch_needed) {
  const int P = B * d;
  std::vector<LooProblem> probs(P);
  std::vector<int2> cta_map;
  std::vector<int> small_ids, tiny_ids;
  const bool tiny_ok = !getenv("CB2_NO_LOO_TINY");
  int maxN = 2;
  long long xs_elems = 0, cp_elems = 0;
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < d; ++k) {
      LooProblem& pr = probs[(size_t)b * d + k];
      pr.x = mu_dev[b] + k; pr.stride = d; pr.N = N[b];
      pr.circular = circular ? circular[k] : 0;
      pr.nrb = (N[b] + kThreads * kRI - 1) / (kThreads * kRI);
      pr.part0 = b * d + k;
      pr.small = N[b] <= kSmallN;
      pr.nstride = (N[b] + 31) / 32 * 32;
      pr.xs_off = xs_elems;
      pr.cp_off = cp_elems;
      xs_elems += pr.nstride;
      cp_elems += (long long)pr.nstride * pr.nrb;
      if (N[b] > maxN) maxN = N[b];
      if (pr.small) (tiny_ok && N[b] <= kTinyN ? tiny_ids : small_ids).push_back(b * d + k);
    }
  // row block rb of a problem visits nrb - rb column tiles: launch the heavy CTAs first (largest-first list scheduling)
  for (int rb = 0; rb < (maxN + kThreads * kRI - 1) / (kThreads * kRI); ++rb)
    for (int p = 0; p < P; ++p)
      if (!probs[p].small && rb < probs[p].nrb) cta_map.push_back(make_int2(p, rb));
  if (maxN > CB2_MAX_PRODUCT_N) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "bandwidth search supports N <= %d (got %d)", CB2_MAX_PRODUCT_N, maxN);
  const bool f64 = ctx->precision == CB2_FP64;
  const size_t tsz = f64 ? sizeof(double) : sizeof(float);
  size_t need = cb2_align(sizeof(LooProblem) * P) + cb2_align(sizeof(LooState) * P) + cb2_align(sizeof(int2) * (cta_map.size() + 1)) + cb2_align(sizeof(int) * (small_ids.size() + tiny_ids.size() + 1)) +
                cb2_align(sizeof(double) * P) + cb2_align(sizeof(int) * 64) + cb2_align(sizeof(double) * (size_t)xs_elems) +
                cb2_align(tsz * (size_t)xs_elems) + cb2_align(tsz * (size_t)cp_elems) + 1024;
  if (scratch_needed) *scratch_needed = need;
  if (!scratch) return CB2_OK;  // sizing query
  if (scratch_bytes < need) return cb2_fail(ctx, CB2_ERR_NOMEM, "loo scratch too small");
  char* base = (char*)scratch;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = base + off; off += cb2_align(bytes); return p; };
  // uploaded tables first and back to back (problems, CTA map, small / tiny ids): one copy
  LooProblem* probs_dev = (LooProblem*)take(sizeof(LooProblem) * P);
  int2* map_dev = (int2*)take(sizeof(int2) * (cta_map.size() + 1));
  int* small_dev = (int*)take(sizeof(int) * (small_ids.size() + tiny_ids.size() + 1));
  int* tiny_dev = small_dev + small_ids.size();
  const size_t tab_bytes = (size_t)((char*)(small_dev + small_ids.size() + tiny_ids.size()) - base);
  LooState* st_dev = (LooState*)take(sizeof(LooState) * P);
  double* partial = (double*)take(sizeof(double) * P);
  int* counters = (int*)take(sizeof(int) * 64);
  double* xs_dev = (double*)take(sizeof(double) * (size_t)xs_elems);
  void* rowsum = take(tsz * (size_t)xs_elems);
  void* colpart = take(tsz * (size_t)cp_elems);
  {
    std::vector<char> tab(tab_bytes);
    memcpy(tab.data() + ((char*)probs_dev - base), probs.data(), sizeof(LooProblem) * P);
    if (!cta_map.empty()) memcpy(tab.data() + ((char*)map_dev - base), cta_map.data(), sizeof(int2) * cta_map.size());
    if (!small_ids.empty()) memcpy(tab.data() + ((char*)small_dev - base), small_ids.data(), sizeof(int) * small_ids.size());
    if (!tiny_ids.empty()) memcpy(tab.data() + ((char*)tiny_dev - base), tiny_ids.data(), sizeof(int) * tiny_ids.size());
    CB2_CUDA(ctx, cudaMemcpyAsync(base, tab.data(), tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  const bool all_tiny = (int)tiny_ids.size() == P;   // then the kernel writes the bandwidths itself and nothing else is launched
  if (!tiny_ids.empty()) {
    if (f64) loo_tiny_kernel<double><<<(int)tiny_ids.size(), kThreads, 0, ctx->stream>>>(probs_dev, tiny_dev, st_dev, all_tiny ? sigma_out_dev : nullptr);
    else loo_tiny_kernel<float><<<(int)tiny_ids.size(), kThreads, 0, ctx->stream>>>(probs_dev, tiny_dev, st_dev, all_tiny ? sigma_out_dev : nullptr);
    CB2_CHECK_LAUNCH(ctx, "loo_tiny_kernel");
    if (all_tiny) return CB2_OK;
  }
  CB2_CUDA(ctx, cudaMemsetAsync(counters, 0, sizeof(int) * 64, ctx->stream));
  if (!small_ids.empty()) {
    // few problems of a few hundred points (one sweep of a small graph): a cluster of CTAs per problem, for latency; many
    // problems: one CTA each with the symmetric pair evaluation, for throughput
    int minN = maxN;
    for (int id : small_ids) minN = std::min(minN, probs[id].N);
    const bool use_cluster = !getenv("CB2_NO_LOO_CLUSTER") && minN >= 192 && (int)small_ids.size() * kLooCluster <= 2 * ctx->sm_count;
    if (use_cluster) {
      if (f64) loo_cluster_kernel<double><<<(int)small_ids.size() * kLooCluster, kLooClusterCta, 0, ctx->stream>>>(probs_dev, small_dev, st_dev);
      else loo_cluster_kernel<float><<<(int)small_ids.size() * kLooCluster, kLooClusterCta, 0, ctx->stream>>>(probs_dev, small_dev, st_dev);
      CB2_CHECK_LAUNCH(ctx, "loo_cluster_kernel");
    } else {
      if (f64) loo_small_kernel<double><<<(int)small_ids.size(), kThreads, 0, ctx->stream>>>(probs_dev, small_dev, st_dev);
      else loo_small_kernel<float><<<(int)small_ids.size(), kThreads, 0, ctx->stream>>>(probs_dev, small_dev, st_dev);
      CB2_CHECK_LAUNCH(ctx, "loo_small_kernel");
    }
  }
  if (!cta_map.empty()) {
  int P2 = 1;
  while (P2 < maxN) P2 <<= 1;
  size_t sort_smem = sizeof(double) * P2;
  cudaError_t e = cb2_func_smem(ctx, (const void*)loo_sort_init_kernel, sort_smem);
  if (e != cudaSuccess) return cb2_fail(ctx, CB2_ERR_CUDA, "loo sort smem attribute: %s", cudaGetErrorString(e));
  loo_sort_init_kernel<<<P, kSortThreads, sort_smem, ctx->stream>>>(probs_dev, st_dev, xs_dev);
  CB2_CHECK_LAUNCH(ctx, "loo_sort_init_kernel");
  const size_t rows_smem = tsz * (size_t)kCT * (1 + kThreads / 32);
  e = f64 ? cb2_func_smem(ctx, (const void*)loo_rows_kernel<double>, rows_smem) : cb2_func_smem(ctx, (const void*)loo_rows_kernel<float>, rows_smem);
  if (e != cudaSuccess) return cb2_fail(ctx, CB2_ERR_CUDA, "loo rows smem attribute: %s", cudaGetErrorString(e));
  const int max_iter = 64, check_every = 8;
  for (int it = 0; it < max_iter; ++it) {
    if (f64) {
      loo_rows_kernel<double><<<(int)cta_map.size(), kThreads, rows_smem, ctx->stream>>>(probs_dev, st_dev, map_dev, xs_dev, (double*)rowsum, (double*)colpart);
      CB2_CHECK_LAUNCH(ctx, "loo_rows_kernel");
      loo_finish_kernel<double><<<P, kThreads, 0, ctx->stream>>>(probs_dev, st_dev, (const double*)rowsum, (const double*)colpart, partial);
    } else {
      loo_rows_kernel<float><<<(int)cta_map.size(), kThreads, rows_smem, ctx->stream>>>(probs_dev, st_dev, map_dev, xs_dev, (float*)rowsum, (float*)colpart);
      CB2_CHECK_LAUNCH(ctx, "loo_rows_kernel");
      loo_finish_kernel<float><<<P, kThreads, 0, ctx->stream>>>(probs_dev, st_dev, (const float*)rowsum, (const float*)colpart, partial);
    }
    CB2_CHECK_LAUNCH(ctx, "loo_finish_kernel");
    golden_step_kernel<<<(P + 127) / 128, 128, 0, ctx->stream>>>(probs_dev, P, partial, st_dev, counters + it);
    CB2_CHECK_LAUNCH(ctx, "golden_step_kernel");
    if ((it + 1) % check_every == 0) {
      int active = 0;
      CB2_CUDA(ctx, cudaMemcpyAsync(&active, counters + it, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (active == 0) break;
    }
  }
  }  // large problems
  loo_collect_kernel<<<(P + 127) / 128, 128, 0, ctx->stream>>>(st_dev, P, sigma_out_dev);
  CB2_CHECK_LAUNCH(ctx, "loo_collect_kernel");
  return CB2_OK;
}

extern "C" int cb2_kde_bw_loo_batch(cb2_ctx* ctx, const double* const* mu, const int32_t* N, int B, int d,
                                    const uint8_t* circular, double* sigma_out) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, mu && N && sigma_out && B > 0, "null pointer or empty batch");
  if (d < 1 || d > CB2_MAX_DIM) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "dimension %d outside 1..%d", d, CB2_MAX_DIM);
  size_t total = 0;
  for (int b = 0; b < B; ++b) {
    CB2_REQUIRE(ctx, mu[b] && N[b] >= 2, "density %d: need at least 2 points", b);
    total += cb2_align((size_t)N[b] * d * sizeof(double));
  }
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t loo_need = 0;
  std::vector<const double*> fake(B, nullptr);
  int rc = cb2_bw_loo_dev_core(ctx, fake.data(), N, B, d, circular, nullptr, nullptr, 0, &loo_need);
  if (rc) return rc;
  size_t out_bytes = cb2_align(sizeof(double) * B * d);
  rc = cb2_arena_reserve(ctx, total + loo_need + out_bytes + 4096);
  if (rc) return rc;
  std::vector<const double*> dev(B);
  // small batches: pack every input into pinned staging and upload with one copy
  const bool packed = total <= ((size_t)4 << 20) && cb2_pinned_reserve(ctx, total + 256) == CB2_OK;
  char* first = nullptr;
  for (int b = 0; b < B; ++b) {
    const size_t nb = (size_t)N[b] * d * sizeof(double);
    double* p = (double*)cb2_arena_alloc(ctx, nb);
    if (!p) return cb2_fail(ctx, CB2_ERR_NOMEM, "loo scratch exhausted");
    if (!first) first = (char*)p;
    if (packed) memcpy(ctx->pinned + ((char*)p - first), mu[b], nb);
    else CB2_CUDA(ctx, cudaMemcpyAsync(p, mu[b], nb, cudaMemcpyHostToDevice, ctx->stream));
    dev[b] = p;
  }
  if (packed) {
    const size_t span = (size_t)((const char*)dev[B - 1] - first) + (size_t)N[B - 1] * d * sizeof(double);
    CB2_CUDA(ctx, cudaMemcpyAsync(first, ctx->pinned, span, cudaMemcpyHostToDevice, ctx->stream));
  }
  double* out_dev = (double*)cb2_arena_alloc(ctx, sizeof(double) * B * d);
  void* scratch = cb2_arena_alloc(ctx, loo_need);
  if (!out_dev || !scratch) return cb2_fail(ctx, CB2_ERR_NOMEM, "loo scratch exhausted");
  rc = cb2_bw_loo_dev_core(ctx, dev.data(), N, B, d, circular, out_dev, scratch, loo_need, nullptr);
  if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
  CB2_CUDA(ctx, cudaMemcpyAsync(sigma_out, out_dev, sizeof(double) * B * d, cudaMemcpyDeviceToHost, ctx->stream));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CB2_OK;
}

extern "C" int cb2_kde_bw_loo(cb2_ctx* ctx, const double* mu, int d, int N, const uint8_t* circular, double* sigma_out) {
  const double* mus[1] = {mu};
  int32_t Ns[1] = {N};
  return cb2_kde_bw_loo_batch(ctx, mus, Ns, 1, d, circular, sigma_out);
}

// Device-resident flavour: mu_dev[b] are device pointers (d x N[b]), sigma_out_dev is B x d on the device.
// For N <= 1024 the whole search of a density runs inside one kernel and nothing is waited for.
extern "C" int cb2_kde_bw_loo_batch_dev(cb2_ctx* ctx, const double* const* mu_dev, const int32_t* N, int B, int d,
                                        const uint8_t* circular, double* sigma_out_dev) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, mu_dev && N && sigma_out_dev && B > 0, "null pointer or empty batch");
  if (d < 1 || d > CB2_MAX_DIM) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "dimension %d outside 1..%d", d, CB2_MAX_DIM);
  for (int b = 0; b < B; ++b) CB2_REQUIRE(ctx, mu_dev[b] && N[b] >= 2, "density %d: need at least 2 points", b);
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t loo_need = 0;
  int rc = cb2_bw_loo_dev_core(ctx, mu_dev, N, B, d, circular, nullptr, nullptr, 0, &loo_need);
  if (rc) return rc;
  rc = cb2_arena_reserve(ctx, loo_need + 4096);
  if (rc) return rc;
  void* scratch = cb2_arena_alloc(ctx, loo_need);
  if (!scratch) return cb2_fail(ctx, CB2_ERR_NOMEM, "loo scratch exhausted");
  return cb2_bw_loo_dev_core(ctx, mu_dev, N, B, d, circular, sigma_out_dev, scratch, loo_need, nullptr);
}
