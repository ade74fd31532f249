// gibbs.cuh -- the multiscale Gibbs product kernel (included by product.cu inside its anonymous namespace).
//
// One thread = one chain; a CTA advances 128 chains (4 warps) in lock step through (level, sweep, message); four CTAs share an SM
// (126 registers per thread, 54 KB of shared memory and 128 tensor-memory columns per CTA).  Per-candidate budget of the loops:
//   leaf level    : FP32 d = 6 wrap-free CTAs score the leaves on the tensor cores (leaf_pass_tc: 3 x TF32 products of per-chain
//                   coefficient rows with per-leaf feature tiles, accumulators in tensor memory); a chain then spends FADD +
//                   MUFU.EX2 + FADD + a share of the running maximum per leaf: bound by the XU pipe (16 MUFU / clk / SM).
//                   Everything else: Euclidean dim = 2 FFMA, circular dim = FFMA + FADD + FMNMX + FFMA, then FADD + EX2 + FADD.
//   coarser levels: node variances differ, so 1 / (var + C) is needed per node.  Packed FP32 d = 6 loops: coordinate pairs in
//                   f32x2 registers (FFMA2 / FMUL2), the two groups of three coordinates over common denominators D_x, D_y,
//                   p = r_x r_y exp2(.. - N_x r_x^2 - N_y r_y^2) with r = rsqrt(D) (one denominator D_x D_y and one rsqrt where a
//                   warp can bound the product, score_node_rsq1): 15 packed + 8 scalar FMA-pipe
//                   instructions, 2 - 3 MUFU and 3 LDS.128 per candidate.  Measured under full load: 60 clk per candidate and
//                   SM sub-partition against 37 clk of FMA pipe and 29 clk-equivalents of the SM-wide shared-memory pipe (a
//                   warp-broadcast LDS.128 costs 2.4 clk of it, bench/micro/lds_bcast.cu) -- both pipes are near their limit there.
// Circular coordinates are stored wrapped to [-pi, pi] (tree_build_kernel) and the chain mean comes out of atan2,
// so |mu - M| <= 2 pi and the wrapped distance is min(|df|, 2 pi - |df|): no rounding instruction needed.
// When every chain of a warp can prove from the message's coordinate range (tree_build_kernel) that no candidate is
// more than pi away from its mean, the warp takes the all-Euclidean loop (`flat`): identical values, 2
// instructions per circular dimension fewer.  Beliefs that do not straddle the +-pi cut always qualify.
// (Two chains per thread -- every record register used twice -- halve the shared-memory loads but leave too few warps to hide
// latency: measured slower, profiles/r01_gibbs_two_chains_experiment.md; the path stays compilable with -DCB2_GIBBS_R=2.)

template <typename T, int N> struct alignas(16) VecRec { T v[N]; };

__device__ unsigned long long g_pick_fallthrough = 0ull;  // pass-2 re-sums that ended short of their target (see chunk_pick)

struct GibbsArgs {
  const DensMeta* metas;
  const ProdDev* prods;
  const Work* work;
  const int* int_pool;
  const void* node_pool;
  uint32_t k0, k1;
  uint32_t circ_mask;
  int level_down;       // 1: weight-proportional random child on the way down, 0: first child
  const double* stat;   // per-message statistics scratch of the tree build; its last 16 doubles = circular ranges
  size_t stat_stride;
};

// what the kernel needs of each message, staged once per CTA in shared memory
template <typename T> struct SMeta {
  uint32_t rec_off32[kMaxLev + 1];  // level record offsets in units of 32 elements (validate_and_plan aligns them)
  uint32_t feat_off32;              // leaf feature tiles of the tensor-core path (units of 32 elements) or 0xffffffff
  uint32_t srec32;                  // SMALL kernels: where this message's records (all levels) sit in shared memory (units of 32 elements)
  long long perm_off;
  T bw2[CB2_MAX_DIM];
  float cmin[CB2_MAX_DIM], cmax[CB2_MAX_DIM];  // range of the stored coordinates (leaves and node means of every level)
  int N, depth, uniform;
  __device__ __forceinline__ size_t rec_off(int l) const { return (size_t)rec_off32[l] * 32; }
};
typedef unsigned short ind_t;  // selected node of a chain on the current level (< 2^14)

template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <> __device__ __forceinline__ float tmin<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double tmin<double>(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float tabs(float a) { return fabsf(a); }
__device__ __forceinline__ double tabs(double a) { return fabs(a); }
__device__ __forceinline__ float tsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double tsqrt(double a) { return sqrt(a); }

// per-chain constants of one (message, level) scoring pass
template <typename T, int DIM> struct ChainConst {
  T a[DIM];  // leaf: sqrt(h_i)            coarse: M_i
  T b[DIM];  // leaf: M_i sqrt(h_i)        coarse: C_i (or 1 when the product has a single message)
  T p[DIM];  // leaf: 2 pi sqrt(h_i)       coarse: unused
  T qs, cmul;
};

// 16-byte vector loads of one padded record; SHARED: 32-bit shared-window address (LDS.128), else read-only global
template <bool SHARED> __device__ __forceinline__ void load16(const float* p, float (&v)[4]) {
  if (SHARED) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(smem_u32(p)));
  else asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
}
template <bool SHARED> __device__ __forceinline__ void load16(const double* p, double (&v)[2]) {
  if (SHARED) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(smem_u32(p)));
  else asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
template <int N, bool SHARED> __device__ __forceinline__ void load_rec(const float* p, float (&v)[N]) {
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    float t[4];
    load16<SHARED>(p + 4 * q, t);
    v[4 * q] = t[0]; v[4 * q + 1] = t[1]; v[4 * q + 2] = t[2]; v[4 * q + 3] = t[3];
  }
}
template <int N, bool SHARED> __device__ __forceinline__ void load_rec(const double* p, double (&v)[N]) {
#pragma unroll
  for (int q = 0; q < N / 2; ++q) {
    double t[2];
    load16<SHARED>(p + 2 * q, t);
    v[2 * q] = t[0]; v[2 * q + 1] = t[1];
  }
}

// chains per thread: a record fetched from shared memory is scored against this many chains
#ifndef CB2_GIBBS_R
#define CB2_GIBBS_R 1
#endif
#ifndef CB2_GIBBS_CTAS
#define CB2_GIBBS_CTAS 4
#endif
template <typename T, int DIM> struct ChainsPerThread { static constexpr int value = CB2_GIBBS_R; };
// bytes of one TMA stage: 8 KB of FP32 records (twice that in FP64), halved when the CTA has half the threads
template <typename T, int DIM> __host__ __device__ constexpr int gibbs_tile_bytes() {
  return kTileBytesF32 * (int)(sizeof(T) / 4) / ChainsPerThread<T, DIM>::value;
}

// dynamic shared memory of the regular kernel: two record stages, chunk sums, selected nodes, barriers, message table (+ the
// tensor-core pass's coefficient rows, MMA barriers and TMEM slot); SMALL kernels append their record staging area behind it
template <typename T, int DIM> __host__ __device__ constexpr size_t gibbs_smem_bytes_c() {
  return 2 * (size_t)gibbs_tile_bytes<T, DIM>() + sizeof(T) * kNCH * kS + sizeof(ind_t) * CB2_MAX_DENS * kS + 16 +
         sizeof(SMeta<T>) * CB2_MAX_DENS + 64 + ((CB2_GIBBS_TC && DIM == 6 && sizeof(T) == 4) ? (size_t)(16384 + 128 + 64) : 0);
}

// ---- packed FP32 pairs (fma.rn.f32x2 -> SASS FFMA2 on sm_100a): one issue slot for two lanes of arithmetic; the
// all-Euclidean scoring loops of six-dimensional products pair the coordinates (0,1) (2,3) (4,5) as LDS.128 delivers them.
template <bool SHARED> __device__ __forceinline__ void load16p(const float* p, f32x2& lo, f32x2& hi) {
  if (SHARED) asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(smem_u32(p)));
  else asm volatile("ld.global.nc.v2.b64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p));
}
template <bool SHARED> __device__ __forceinline__ void load16p(const double*, f32x2&, f32x2&) {}

// one candidate record in registers: N values of T, or (PACK) N/2 packed pairs
template <typename T, int N, bool PACK> struct RecR {
  T v[PACK ? 1 : N];
  f32x2 q[PACK ? N / 2 : 1];
};
template <bool SHARED, typename T, int N, bool PACK> __device__ __forceinline__ void load_recr(const T* p, RecR<T, N, PACK>& r) {
  if constexpr (PACK) {
#pragma unroll
    for (int k = 0; k < N / 4; ++k) load16p<SHARED>(p + 4 * k, r.q[2 * k], r.q[2 * k + 1]);
  } else {
    load_rec<N, SHARED>(p, r.v);
  }
}

// per-chain constants of the packed loops (six dimensions as three pairs)
struct PackedConst {
  f32x2 a[3];  // leaf: sqrt(h)           coarse: -M (CB2_COARSE_RSQ: -M sqrt(qs))
  f32x2 b[3];  // leaf: -M sqrt(h)        coarse: C (1 for a single message)
  f32x2 cmul;
  float nqs;   // -qs                     (CB2_COARSE_RSQ: sqrt(qs), the scale of the coordinate differences)
};
template <typename T, int DIM> __device__ __forceinline__ PackedConst make_packed(const ChainConst<T, DIM>& c, bool leaf) {
  PackedConst p = {};
  if constexpr (DIM == 6 && sizeof(T) == 4) {
#if CB2_COARSE_RSQ
    const float sq = leaf ? 1.0f : sqrtf((float)c.qs);
    p.nqs = sq;
#else
    const float sq = 1.0f;
    p.nqs = -c.qs;
#endif
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      p.a[k] = leaf ? pack2(c.a[2 * k], c.a[2 * k + 1]) : pack2(-c.a[2 * k] * sq, -c.a[2 * k + 1] * sq);
      p.b[k] = leaf ? pack2(-c.b[2 * k], -c.b[2 * k + 1]) : pack2(c.b[2 * k], c.b[2 * k + 1]);
    }
    p.cmul = pack2(c.cmul, c.cmul);
  }
  return p;
}

// log2-score of a leaf relative to the running reference m (nm = -m):
//   log2 w - m - sum_i wrapdiff_i^2 * h_i   (h_i = 0.5 log2e / (bw_i^2 + C_i); the log-variance term is the same
//   for every leaf of the level and cancels in the normalisation; with uniform weights so does log2 w: UW)
template <typename T, int DIM, bool UW, int RECL>
__device__ __forceinline__ T score_leaf(const T (&r)[RECL], const ChainConst<T, DIM>& c, uint32_t circ, T nm) {
  T e = UW ? nm : r[DIM] + nm;
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    T df = fma(r[i], c.a[i], -c.b[i]);
    if (circ >> i & 1) {
      T ad = tabs(df);
      df = tmin(ad, c.p[i] - ad);
    }
    e = fma(-df, df, e);
  }
  return e;
}

// log2-score of a coarse node relative to m: log2 w - m - qs sum_i d_i^2/(var_i + C_i) - 0.5 sum_i log2(var_i + C_i)
template <typename T, int DIM, int RECN>
__device__ __forceinline__ T score_node(const T (&r)[RECN], const ChainConst<T, DIM>& c, uint32_t circ, T nm) {
  T t[DIM], sv[DIM];
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    T df = r[i] - c.a[i];
    if (circ >> i & 1) {
      T ad = tabs(df);
      df = tmin(ad, T(CB2_TWO_PI) - ad);
    }
    t[i] = df * df;
    sv[i] = fma(r[DIM + i], c.cmul, c.b[i]);
  }
  T e = r[2 * DIM] + nm;
#pragma unroll
  for (int g = 0; g < DIM; g += 3) {
    if (g + 2 < DIM) {  // three dims over one denominator
      T s01 = sv[g] * sv[g + 1], s12 = sv[g + 1] * sv[g + 2], s02 = sv[g] * sv[g + 2];
      T num = fma(t[g], s12, fma(t[g + 1], s02, t[g + 2] * s01));
      T den = s01 * sv[g + 2];
      e = fma(-c.qs * num, fast_rcp(den), e);
      e = fma(T(-0.5), fast_log2(den), e);
    } else if (g + 1 < DIM) {
      T den = sv[g] * sv[g + 1];
      T num = fma(t[g], sv[g + 1], t[g + 1] * sv[g]);
      e = fma(-c.qs * num, fast_rcp(den), e);
      e = fma(T(-0.5), fast_log2(den), e);
    } else {
      e = fma(-c.qs * t[g], fast_rcp(sv[g]), e);
      e = fma(T(-0.5), fast_log2(sv[g]), e);
    }
  }
  return e;
}

// packed leaf score: same quantity as score_leaf without wrap; the squares are summed as (d0^2+d2^2+d4^2) + (d1^2+d3^2+d5^2)
template <bool UW>
__device__ __forceinline__ float score_leaf_p(const f32x2 (&q)[4], const PackedConst& c, float nm) {
  const f32x2 d0 = fma2(q[0], c.a[0], c.b[0]), d1 = fma2(q[1], c.a[1], c.b[1]), d2 = fma2(q[2], c.a[2], c.b[2]);
  // the reference -m rides in the first lane of the accumulator: s = (sum_even - nm, sum_odd), e = -s.x - s.y is one FADD
  f32x2 s = fma2(d0, d0, pack2(-nm, 0.0f));
  s = fma2(d1, d1, s);
  s = fma2(d2, d2, s);
  float sx, sy, lw, pad;
  unpack2(s, sx, sy);
  unpack2(q[3], lw, pad);
  return UW ? -sx - sy : (lw - sx) - sy;
}

// packed coarse-node score: the two common-denominator groups are the even and the odd coordinates
__device__ __forceinline__ float score_node_p(const f32x2 (&q)[8], const PackedConst& c, float nm) {
  const f32x2 d0 = add2(q[0], c.a[0]), d1 = add2(q[1], c.a[1]), d2 = add2(q[2], c.a[2]);
  const f32x2 t0 = mul2(d0, d0), t1 = mul2(d1, d1), t2 = mul2(d2, d2);
  const f32x2 s0 = fma2(q[3], c.cmul, c.b[0]), s1 = fma2(q[4], c.cmul, c.b[1]), s2 = fma2(q[5], c.cmul, c.b[2]);
  const f32x2 s01 = mul2(s0, s1), s12 = mul2(s1, s2), s02 = mul2(s0, s2);
  const f32x2 num = fma2(t0, s12, fma2(t1, s02, mul2(t2, s01)));
  const f32x2 den = mul2(s01, s2);
  float nx, ny, dx, dy, lw, p1;
  unpack2(num, nx, ny);
  unpack2(den, dx, dy);
  unpack2(q[6], lw, p1);
  float e = lw + nm;
  e = fma(c.nqs * nx, fast_rcp(dx), e);
  e = fma(-0.5f, fast_log2(dx), e);
  e = fma(c.nqs * ny, fast_rcp(dy), e);
  e = fma(-0.5f, fast_log2(dy), e);
  return e;
}

// CB2_COARSE_RSQ: the same coarse-node probability with three MUFU instead of five.  With D_g = prod (var_i + C_i) over the
// coordinates of group g (even / odd) and N_g = sum_i qs d_i^2 prod_{k != i} (var_k + C_k):
//   p = w 2^{-m} prod_g D_g^{-1/2} exp2(-N_g / D_g) = (r_x r_y) exp2(lw - m - N_x r_x^2 - N_y r_y^2),   r_g = rsqrt(D_g):
// one MUFU.RSQ per group (its square is the reciprocal) and one MUFU.EX2, against 2 RCP + 2 LG2 + 1 EX2 when the log-variance
// term is carried inside the exponent.  Returns the exponent e' (relative to the reference m) and the multiplier r = r_x r_y;
// the caller sums r exp2(e').  The reference m of a pass is the FULL log2-score of the level's first candidate (tile_pass),
// so r exp2(e') = 2^{score - m} exactly as before and only the bookkeeping of the lazily raised maximum differs: it compares
// e' with 60 - log2 r_0 (r_0 of the first candidate), which is the old test for candidates whose variances resemble the first
// candidate's and keeps r exp2(e') <= 2^60 r / r_0 in every case.
__device__ __forceinline__ float fast_rsq(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// UW: every node of the level carries the same weight (equal kernel weights and N a power of two): log2 w cancels in the
// normalisation and its load (the fourth LDS of the record: the shared-memory pipe is as busy as the FMA pipe here) is dropped
template <bool UW>
__device__ __forceinline__ float score_node_rsq(const f32x2 (&q)[8], const PackedConst& c, float nm, float& mul) {
  const f32x2 sq = pack2(c.nqs, c.nqs);
  const f32x2 d0 = fma2(q[0], sq, c.a[0]), d1 = fma2(q[1], sq, c.a[1]), d2 = fma2(q[2], sq, c.a[2]);
  const f32x2 t0 = mul2(d0, d0), t1 = mul2(d1, d1), t2 = mul2(d2, d2);
  const f32x2 s0 = fma2(q[3], c.cmul, c.b[0]), s1 = fma2(q[4], c.cmul, c.b[1]), s2 = fma2(q[5], c.cmul, c.b[2]);
  const f32x2 s01 = mul2(s0, s1);
  const f32x2 num = fma2(s2, fma2(t1, s0, mul2(t0, s1)), mul2(s01, t2));
  const f32x2 den = mul2(s01, s2);
  float nx, ny, dx, dy, lw, p1;
  unpack2(num, nx, ny);
  unpack2(den, dx, dy);
  unpack2(q[6], lw, p1);
  const float rx = fast_rsq(dx), ry = fast_rsq(dy);
  float e = UW ? nm : lw + nm;
  e = fma(-nx, rx * rx, e);
  e = fma(-ny, ry * ry, e);
  mul = rx * ry;
  return e;
}

// One common denominator over all six coordinates: D = D_x D_y, N = N_x D_y + N_y D_x, p = r exp2(lw - m - N r^2), r = rsqrt(D):
// two MUFU per candidate instead of three for the same FMA-pipe work.  The six-fold product of variances squares the dynamic range,
// so a warp takes this form only when every one of its chains can prove from the message's bounds (kernel variance below, half
// the coordinate range squared + kernel variance above) that D stays inside [1e-30, 1e30] for every node: gibbs_kernel, `one_den`.
template <bool UW>
__device__ __forceinline__ float score_node_rsq1(const f32x2 (&q)[8], const PackedConst& c, float nm, float& mul) {
  const f32x2 sq = pack2(c.nqs, c.nqs);
  const f32x2 d0 = fma2(q[0], sq, c.a[0]), d1 = fma2(q[1], sq, c.a[1]), d2 = fma2(q[2], sq, c.a[2]);
  const f32x2 t0 = mul2(d0, d0), t1 = mul2(d1, d1), t2 = mul2(d2, d2);
  const f32x2 s0 = fma2(q[3], c.cmul, c.b[0]), s1 = fma2(q[4], c.cmul, c.b[1]), s2 = fma2(q[5], c.cmul, c.b[2]);
  const f32x2 s01 = mul2(s0, s1);
  const f32x2 num = fma2(s2, fma2(t1, s0, mul2(t0, s1)), mul2(s01, t2));
  const f32x2 den = mul2(s01, s2);
  float nx, ny, dx, dy, lw, p1;
  unpack2(num, nx, ny);
  unpack2(den, dx, dy);
  unpack2(q[6], lw, p1);
  const float r = fast_rsq(dx * dy);
  const float n6 = fmaf(nx, dy, ny * dx);
  mul = r;
  return fmaf(-n6, r * r, UW ? nm : lw + nm);
}

// MODE 0: coarse level, 1: leaf level with per-kernel weights, 2: leaf level with uniform weights, 3: coarse level whose nodes
// all weigh the same (packed loops only); 4 / 5: modes 0 / 3 over one common denominator (packed loops only)
template <int MODE> struct IsCoarse { static constexpr bool value = MODE == 0 || MODE == 3 || MODE == 4 || MODE == 5; };
template <int DIM, int MODE> struct RecLen { static constexpr int value = IsCoarse<MODE>::value ? (2 * DIM + 1 + 3) / 4 * 4 : (DIM + 1 + 3) / 4 * 4; };

// whether a scoring path returns the probability as mul * exp2(e) (CB2_COARSE_RSQ) instead of exp2(e)
template <int MODE, bool PACK> struct HasMul { static constexpr bool value = CB2_COARSE_RSQ && PACK && IsCoarse<MODE>::value; };

template <typename T, int DIM, int MODE, bool PACK>
__device__ __forceinline__ T score_regs(const RecR<T, RecLen<DIM, MODE>::value, PACK>& r, const ChainConst<T, DIM>& c,
                                        const PackedConst& pc, uint32_t circ, T nm, T& mul) {
  mul = T(1);
  if constexpr (PACK) {
    if constexpr (MODE == 4 || MODE == 5) return score_node_rsq1<MODE == 5>(r.q, pc, nm, mul);
    else if constexpr (MODE == 0 || MODE == 3) {
#if CB2_COARSE_RSQ
      return score_node_rsq<MODE == 3>(r.q, pc, nm, mul);
#else
      return score_node_p(r.q, pc, nm);
#endif
    }
    else return score_leaf_p<MODE == 2>(r.q, pc, nm);
  } else {
    if constexpr (IsCoarse<MODE>::value) return score_node<T, DIM>(r.v, c, circ, nm);
    else return score_leaf<T, DIM, MODE == 2>(r.v, c, circ, nm);
  }
}

// one scoring pass over the candidates of a level held in a shared-memory tile: chunk partial sums of exp2(e - m) for
// the R chains of this thread; every record is fetched once and scored against all R chains
// (a branch-free variant -- sum the tile optimistically, test the largest exponent once per tile, repeat the tile after a raise --
// was measured slower here: 213 vs 192 ms per 256 variables, and so was the same scheme per group of eight, 178.5 vs 171.9: with
// the exponentials taken one by one behind each score the MUFU results arrive late; the tensor-core leaf pass below does use that
// scheme per tile)
template <typename T, int DIM, int MODE, bool PACK, int R>
__device__ __forceinline__ void tile_pass(const T* __restrict__ tb, int nt, int zbase, int CH, const ChainConst<T, DIM> (&cc)[R],
                                          const PackedConst (&pc)[R], uint32_t circ, T* __restrict__ chunk, int tid, T (&m)[R],
                                          T (&csum)[R], int& cidx, T (&thr)[R]) {
  constexpr int REC = RecLen<DIM, MODE>::value;
  constexpr int NT = kS / R;   // threads per CTA; chain r of this thread is column r * NT + tid of the per-chain arrays
  constexpr int G = kG / R;    // candidates scored together: G * R independent evaluations in flight
  const int ntg = nt & ~(kG - 1);
  if (zbase == 0) {  // the reference starts at the score of the level's first candidate (any finite score works; it is raised lazily)
    RecR<T, REC, PACK> rec0;
    load_recr<true>(tb, rec0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      T mul0;
      m[r] = score_regs<T, DIM, MODE, PACK>(rec0, cc[r], pc[r], circ, T(0), mul0);
      if constexpr (HasMul<MODE, PACK>::value) {  // m = full score of the first candidate; the raise test moves by log2 r_0
        const T l0 = fast_log2(mul0);
        m[r] += l0;
        thr[r] = T(60) - l0;
      } else {
        thr[r] = T(60);
      }
    }
  }
  T nm[R];
#pragma unroll
  for (int r = 0; r < R; ++r) nm[r] = -m[r];
  for (int z = 0; z < ntg; z += G) {
    T e[R][G], mu[R][G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      RecR<T, REC, PACK> rec;
      load_recr<true>(tb + (z + g) * REC, rec);
#pragma unroll
      for (int r = 0; r < R; ++r) e[r][g] = score_regs<T, DIM, MODE, PACK>(rec, cc[r], pc[r], circ, nm[r], mu[r][g]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      T gm = e[r][0];
#pragma unroll
      for (int g = 1; g < G; ++g) gm = fmax(gm, e[r][g]);
      if (gm > thr[r]) {  // rare: raise the reference and rescale what has been summed so far
        gm -= thr[r] - T(60);  // (HasMul: the exponents sit log2 r_0 below the scores)
        T sc = fast_exp2(-gm);
        csum[r] *= sc;
        for (int c = 0; c < cidx; ++c) chunk[c * kS + r * NT + tid] *= sc;
        m[r] += gm;
#pragma unroll
        for (int g = 0; g < G; ++g) e[r][g] -= gm;
        nm[r] = -m[r];
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if constexpr (HasMul<MODE, PACK>::value) csum[r] = fma(fast_exp2(e[r][g]), mu[r][g], csum[r]);
        else csum[r] += fast_exp2(e[r][g]);
      }
    }
    if (((zbase + z + G) & (CH - 1)) == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) { chunk[cidx * kS + r * NT + tid] = csum[r]; csum[r] = T(0); }
      ++cidx;
    }
  }
  for (int z = ntg; z < nt; ++z) {  // ragged tail of the level (count not a multiple of kG)
    RecR<T, REC, PACK> rec;
    load_recr<true>(tb + z * REC, rec);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      T mul;
      T e = score_regs<T, DIM, MODE, PACK>(rec, cc[r], pc[r], circ, nm[r], mul);
      if (e > thr[r]) {
        const T up = e - (thr[r] - T(60));
        T sc = fast_exp2(-up);
        csum[r] *= sc;
        for (int c = 0; c < cidx; ++c) chunk[c * kS + r * NT + tid] *= sc;
        m[r] += up;
        e -= up;
        nm[r] = -m[r];
      }
      csum[r] += fast_exp2(e) * mul;
    }
  }
}

// pass 2: re-score the selected chunk from global memory until the cumulative sum crosses the target
template <typename T, int DIM, int MODE, bool PACK, bool SHARED = false>
__device__ __forceinline__ int chunk_pick(const T* __restrict__ lrec, int z0, int z1, const ChainConst<T, DIM>& cc,
                                          const PackedConst& pc, uint32_t circ, T m, T cum, T tgt) {
  constexpr int REC = RecLen<DIM, MODE>::value;
  int pick = -1, lastpos = z0;
  const T nm = -m;
  // PR candidates per round: their (divergent, L2-resident) record loads are in flight together (8 per round measured no
  // faster: other warps cover the latency); the cumulative sum is still taken strictly in candidate order, so the draw is the same as a
  // one-by-one scan
  constexpr int PR = 4;
  for (int z = z0; z < z1 && pick < 0; z += PR) {
    T p[PR];
#pragma unroll
    for (int g = 0; g < PR; ++g) {
      const int zz = min(z + g, z1 - 1);
      RecR<T, REC, PACK> rec;
      load_recr<SHARED>(lrec + (size_t)zz * REC, rec);
      T mul;
      const T e = score_regs<T, DIM, MODE, PACK>(rec, cc, pc, circ, nm, mul);
      p[g] = fast_exp2(e) * mul;
    }
#pragma unroll
    for (int g = 0; g < PR; ++g) {
      if (pick < 0 && z + g < z1) {
        cum += p[g];
        if (p[g] > T(0)) lastpos = z + g;
        if (cum >= tgt) pick = z + g;
      }
    }
  }
  // The cumulative re-sum fell short of the target: pass 1 and pass 2 round differently (chunk sums are added in another
  // order; with the tensor-core leaf pass the pass-1 scores come from a 3 x TF32 product, pass 2 from FP32 FMAs, equal to ~1e-6).
  // The draw then takes the last candidate with mass in the chunk -- the neighbour of the exact answer.  Counted, so that the
  // tests can bound how often it happens (cb2_debug_counter).
  if (pick < 0) atomicAdd(&g_pick_fallthrough, 1ull);
  return pick < 0 ? lastpos : pick;
}

// Gaussian product of the currently selected nodes of all messages except `skip` (skip = -1: all)
template <typename T, int DIM, bool SMALL = false>
__device__ __forceinline__ void combine(const T* __restrict__ pool, const SMeta<T>* __restrict__ sm, int D,
                                        const ind_t* __restrict__ ind_s, int tid, int step, int skip, uint32_t circ,
                                        T (&M)[DIM], T (&C)[DIM]) {
  constexpr int RECN = (2 * DIM + 1 + 3) / 4 * 4, RECL = (DIM + 1 + 3) / 4 * 4;
  T lam[DIM], s1[DIM], s2[DIM];
#pragma unroll
  for (int i = 0; i < DIM; ++i) { lam[i] = T(0); s1[i] = T(0); s2[i] = T(0); }
  const int nk = skip >= 0 ? D - 1 : D;
#pragma unroll 4
  for (int kk = 0; kk < nk; ++kk) {  // branch-free skip so that the independent gathers of several messages overlap
    const int k = kk + ((skip >= 0 && kk >= skip) ? 1 : 0);
    const SMeta<T>& dm = sm[k];
    const int lev = min(step, dm.depth);
    const int node = ind_s[k * kS + tid];
    T mu[DIM], var[DIM];
    // SMALL: `pool` is the shared-memory staging area and the message's levels sit at srec32 in the order of the node pool
    const size_t lbase = SMALL ? (size_t)dm.srec32 * 32 + (dm.rec_off(lev) - dm.rec_off(0)) : dm.rec_off(lev);
    if (lev == dm.depth) {
      const VecRec<T, RECL> r = *reinterpret_cast<const VecRec<T, RECL>*>(pool + lbase + (size_t)node * RECL);
#pragma unroll
      for (int i = 0; i < DIM; ++i) { mu[i] = r.v[i]; var[i] = dm.bw2[i]; }
    } else {
      const VecRec<T, RECN> r = *reinterpret_cast<const VecRec<T, RECN>*>(pool + lbase + (size_t)node * RECN);
#pragma unroll
      for (int i = 0; i < DIM; ++i) { mu[i] = r.v[i]; var[i] = r.v[DIM + i]; }
    }
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      T l = fast_rcp(var[i]);
      lam[i] += l;
      if (circ >> i & 1) {
        T sn, cs;
        sincos_t(mu[i], &sn, &cs);
        s1[i] = fma(l, sn, s1[i]); s2[i] = fma(l, cs, s2[i]);
      } else {
        s1[i] = fma(l, mu[i], s1[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    C[i] = fast_rcp(lam[i]);
    M[i] = (circ >> i & 1) ? atan2_t(s1[i], s2[i]) : s1[i] * C[i];
  }
}

// ---- leaf level on the tensor cores (FP32, d = 6, CTA-wide wrap-free passes; compiled in with -DCB2_GIBBS_TC=1) ----------------
// The leaf score is bilinear in per-chain coefficients and per-leaf features:
//   log2 w - sum_i (a_i x_i - b_i)^2 = [x_i^2 (6), x_i (6), 1, log2 w, 0, 0] . [-a_i^2 (6), 2 a_i b_i (6), -sum b_i^2, 1, 0, 0],
// a 128-chain x kTcTile-leaf x 16 product per feature tile (kTcTile = 128: one accumulator of 128 tensor-memory columns per CTA).
// Both operands are split into TF32 hi + lo parts and three products (hi.hi + hi.lo + lo.hi) are accumulated in FP32 in tensor
// memory: tcgen05.mma kind::tf32, M = 128, N = kTcTile, K = 8, six instructions per tile; operands K-major without swizzle
// (8 x 16-byte core matrices, LBO = 128 B between the K halves, SBO = 512 B between 8-row groups).  The feature tiles come from the
// tree build (product.cu, kTcTile) by the same 1-D bulk copies as the records; every thread writes its chain's coefficient row
// into shared memory once per pass.  Each chain then reads ITS row of the accumulator (TMEM lane = thread) with tcgen05.ld, 32
// leaves at a time.  What the measurements say about this pass (bench/micro/tcgen05_issue_probe.cu, -DCB2_TC_PROF):
//   * a tcgen05.mma of this shape keeps its ISSUING warp for ~75 clk (N = 64 and N = 128 alike; 150 clk when the descriptors
//     are not warp-uniform), dependent or not: the six MMAs of a tile cost the issuing warp ~540 clk and complete ~70 clk later;
//   * with two accumulator buffers of 64 leaves the MMA of the next tile does overlap the sums of this one, but the issuing warp
//     is itself a consumer, so every tile still costs (sums + issue) on that warp and the other warps wait for it: measured no
//     faster than one buffer of 128 leaves, which halves the MMA instructions per leaf (195.9 vs 192.4 ms per 256 variables);
//     a dedicated issuing warp would need a fifth warp of registers that the four CTAs per SM do not have;
//   * under full load the exposed MMA latency is covered by the other three CTAs of the SM: the pass runs at ~70 % of the XU limit.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
// Executed by a CONVERGED warp, one elected lane issues: the operands stay warp-uniform, so the descriptors live in uniform
// registers.  Issued from a divergent `tid == 0` branch every MMA pays register-to-uniform moves and costs the issuing thread
// ~150 clk instead of ~75 (bench/micro/tcgen05_issue_probe.cu: chain of six 906 vs 541 clk).
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct TcState {
  float* sA;          // [2][128 x 16] coefficient rows, hi then lo (canonical operand layout)
  uint64_t* mbars;    // [0..1] accumulator-full barriers (MMA commit), [2..3] accumulator-free barriers (one arrival per warp)
  uint32_t tmem;      // base of this CTA's 128 accumulator columns (kTcBufs buffers of kTcTile columns)
  uint32_t mu0, mu1;  // completed phases of the two accumulator-full barriers
  uint32_t fu0, fu1;  // completed phases of the two accumulator-free barriers
};
constexpr int kTcBufs = 128 / kTcTile;  // accumulator buffers (and feature stages) per CTA: 2 x 64 leaves or 1 x 128
static_assert(kTcTile == 64 || kTcTile == 128, "feature tiles hold 64 or 128 leaves");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// Pipeline of one leaf pass (tiles t = 0 .. ntl-1 of kTcTile leaves, buffer b = t mod kTcBufs):
//   thread 0   : bulk copy of tile t into stage b (as soon as MMA(t - kTcBufs) has completed), six MMAs of tile t into accumulator b
//                as soon as every warp has released it (accumulator-free barrier, one arrival per warp), commit -> full barrier b
//   every warp : waits for full barrier b, reads its 32 rows of accumulator b in blocks of 32 leaves (tcgen05.ld), sums
//                exp2(score - m) into the chain's chunk sums, releases the accumulator.
// With two buffers the tensor core works on tile t + 1 while tile t is summed and no CTA-wide barrier sits inside the pass; the
// warps drift by at most one tile.  The sums are taken optimistically (see tile_pass): no branch inside a block of 32 leaves, the
// largest exponent of the tile is checked once per tile, and a warp with a chain above the threshold restores, raises and
// re-reads the tile (its accumulator has not been released yet).
template <typename T, int DIM>
__device__ __forceinline__ void leaf_pass_tc(const float* __restrict__ feat, int cnt, int CH, const ChainConst<T, DIM>& cc, bool uniform,
                                             unsigned char* smem_tiles, int tile_bytes, uint64_t* bars, uint32_t& use0, uint32_t& use1,
                                             TcState& tc, float* __restrict__ chunk, int tid, float& m, float& csum, int& cidx,
                                             const float* __restrict__ cur_rec) {
  const int ntl = (cnt + kTcTile - 1) / kTcTile;
  auto load_tile = [&](int t) {  // thread 0: one bulk copy (hi + lo parts of kTcTile leaves)
    const int b = t % kTcBufs;
    mbar_expect_tx(&bars[b], (uint32_t)(kTcTileFloats * 4));
    tma_load_1d(smem_tiles + b * tile_bytes, feat + (size_t)t * kTcTileFloats, (uint32_t)(kTcTileFloats * 4), &bars[b]);
  };
  if (tid == 0) {  // the first feature tiles travel while the coefficient rows are written
    load_tile(0);
    if (kTcBufs > 1 && ntl > 1) load_tile(1);
  }
  // ---- this chain's coefficient row, split hi / lo, into the A operand
  {
    float u[16];
    float sb = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float a = (float)cc.a[i], b = (float)cc.b[i];
      u[i] = -a * a; u[6 + i] = 2.f * a * b; sb = fmaf(b, b, sb);
    }
    // with uniform weights log2 w is the same for every leaf and is dropped, exactly as score_leaf<UW> does (pass 2 must see
    // the same reference); the padding leaves of the last tile are masked below instead of relying on their log2 w = -1e30
    // The reference of the pass is the score of the leaf this chain holds at the moment (a likely one), and it rides in the
    // constant coefficient: the accumulator then IS score - m and the sums below need no subtraction per leaf (one FMA-pipe
    // instruction less on the pipe the coarse levels of the other CTAs are waiting for).  A later raise of the reference is
    // carried separately (`extra`).
    {
      float e0 = uniform ? 0.f : cur_rec[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float df = fmaf(cur_rec[i], (float)cc.a[i], -(float)cc.b[i]);
        e0 = fmaf(-df, df, e0);
      }
      m = e0;
    }
    u[12] = -sb - m; u[13] = uniform ? 0.f : 1.f; u[14] = 0.f; u[15] = 0.f;
    float* rowp = tc.sA + (tid >> 3) * 128 + (tid & 7) * 4;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float hi = tf32_round(u[k]);
      rowp[(k >> 2) * 32 + (k & 3)] = hi;
      rowp[2048 + (k >> 2) * 32 + (k & 3)] = tf32_round(u[k] - hi);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcTile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t sA_hi = smem_u32(tc.sA), sA_lo = sA_hi + 8192;
  auto issue_mma = [&](int t) {  // warp 0 (converged): wait for the tile's bytes, six MMAs into accumulator t mod kTcBufs, commit
    const int b = t % kTcBufs;
    mbar_wait(&bars[b], ((b ? use1 : use0) + (uint32_t)(t / kTcBufs)) & 1);  // use count of stage b within this pass
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t sB_hi = smem_u32(smem_tiles + b * tile_bytes), sB_lo = sB_hi + kTcTile * 64;
    const uint32_t d = tc.tmem + (uint32_t)(b * kTcTile);
    umma_tf32(d, umma_desc(sA_hi, 128, 512), umma_desc(sB_hi, 128, 512), idesc, 0u);
    umma_tf32(d, umma_desc(sA_hi + 256, 128, 512), umma_desc(sB_hi + 256, 128, 512), idesc, 1u);
    umma_tf32(d, umma_desc(sA_hi, 128, 512), umma_desc(sB_lo, 128, 512), idesc, 1u);
    umma_tf32(d, umma_desc(sA_hi + 256, 128, 512), umma_desc(sB_lo + 256, 128, 512), idesc, 1u);
    umma_tf32(d, umma_desc(sA_lo, 128, 512), umma_desc(sB_hi, 128, 512), idesc, 1u);
    umma_tf32(d, umma_desc(sA_lo + 256, 128, 512), umma_desc(sB_hi + 256, 128, 512), idesc, 1u);
    umma_commit(&tc.mbars[b]);
  };
  if (tid < 32) {  // every accumulator is free at the start of a pass
    issue_mma(0);
    if (kTcBufs > 1 && ntl > 1) issue_mma(1);
  }
  const uint32_t lane_base = (uint32_t)((tid >> 5) * 32) << 16;
  float extra = 0.f;  // how far this chain has raised its reference above the one folded into the coefficient row
#ifdef CB2_TC_PROF
  long long pw = 0, pc_ = 0, ps = 0, p0, pld = 0, pfree = 0, pissue = 0;
#endif
  for (int t = 0; t < ntl; ++t) {
#ifdef CB2_TC_PROF
    p0 = clock64();
#endif
    const int b = t % kTcBufs;
    mbar_wait(&tc.mbars[b], (b ? tc.mu1 : tc.mu0) & 1);
    if (b) ++tc.mu1; else ++tc.mu0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef CB2_TC_PROF
    pw += clock64() - p0; p0 = clock64();
#endif
    if (tid == 0 && t + kTcBufs < ntl) load_tile(t + kTcBufs);   // MMA(t) is complete: its stage may be refilled
    const float csum0 = csum;
    const int cidx0 = cidx;
    for (;;) {
      float gmax = -INFINITY;
      const bool shifted = __any_sync(0xffffffffu, extra != 0.f);  // some chain of the warp has raised its reference in this pass
      const float nx = -extra;
      // (loading the next block's accumulators while this one is summed -- two blocks of 32 or of 16 registers -- was measured
      // slower, 181.5 / 178.4 vs 167.7 ms: the other warps of the SM already cover the load latency, the extra registers spill)
#pragma unroll 1  // one copy of the 32-leaf body: the kernel is large enough to miss in the instruction cache
      for (int c0 = 0; c0 < kTcTile; c0 += 32) {
        const int z0 = t * kTcTile + c0;
        if (z0 >= cnt) break;  // only padding leaves left
        uint32_t r[32];
#ifdef CB2_TC_PROF
        long long q0 = clock64();
#endif
        tmem_ld32(tc.tmem + lane_base + (uint32_t)(b * kTcTile + c0), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#ifdef CB2_TC_PROF
        pld += clock64() - q0;
#endif
        if (z0 + 32 > cnt) {  // the block that straddles the end of the level: padding leaves carry no mass
#pragma unroll
          for (int g = 0; g < 32; ++g)
            if (z0 + g >= cnt) r[g] = 0xff800000u;  // -inf
        }
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        if (!shifted) {
#pragma unroll
          for (int g = 0; g < 32; ++g) {
            const float e = __uint_as_float(r[g]);
            gmax = fmaxf(gmax, e);
            s[g & 3] += fast_exp2(e);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 32; ++g) {
            const float e = __uint_as_float(r[g]) + nx;
            gmax = fmaxf(gmax, e);
            s[g & 3] += fast_exp2(e);
          }
        }
        csum += (s[0] + s[1]) + (s[2] + s[3]);
        if (((z0 + 32) & (CH - 1)) == 0 && z0 + 32 <= cnt) { chunk[cidx * kS + tid] = csum; csum = 0.f; ++cidx; }
      }
      // the sync.aligned accumulator loads need the whole warp: a warp repeats the tile if any of its chains is above the threshold
      if (!__any_sync(0xffffffffu, gmax > 60.f)) break;
      if (gmax > 60.f) {
        const float up = gmax, sc = fast_exp2(-up);
        m += up;
        extra += up;
        csum = csum0 * sc;
        for (int c = 0; c < cidx0; ++c) chunk[c * kS + tid] *= sc;
      } else {
        csum = csum0;
      }
      cidx = cidx0;
    }
#ifdef CB2_TC_PROF
    pc_ += clock64() - p0; p0 = clock64();
#endif
    if (t + kTcBufs < ntl) {  // release accumulator b; thread 0 refills it once every warp has
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&tc.mbars[2 + b]);
      if (tid < 32) {
#ifdef CB2_TC_PROF
        long long q0 = clock64();
#endif
        mbar_wait(&tc.mbars[2 + b], ((b ? tc.fu1 : tc.fu0) + (uint32_t)(t / kTcBufs)) & 1);
#ifdef CB2_TC_PROF
        pfree += clock64() - q0; q0 = clock64();
#endif
        issue_mma(t + kTcBufs);
#ifdef CB2_TC_PROF
        pissue += clock64() - q0;
#endif
      }
    }
#ifdef CB2_TC_PROF
    ps += clock64() - p0;
#endif
  }
#ifdef CB2_TC_PROF
  if (blockIdx.x == 0 && (tid == 32 || tid == 0) && ntl >= 32) printf("tc prof tid %d tiles %d (clk/tile): wait-mma %lld ld+compute %lld (ld latency %lld) release+issue %lld (wait-free %lld tma-wait+issue %lld)\n", tid, ntl, pw / ntl, pc_ / ntl, pld / ntl, ps / ntl, pfree / ntl, pissue / ntl);
#endif
  // every thread keeps the phase counters in step (only thread 0 waited on the copy and the accumulator-free barriers)
  const int nrel0 = ntl > kTcBufs ? (ntl - kTcBufs + (kTcBufs - 1)) / kTcBufs : 0;  // releases of buffer 0: tiles t = 0, kTcBufs, .. with t + kTcBufs < ntl
  const int nrel1 = (kTcBufs > 1 && ntl > kTcBufs) ? (ntl - kTcBufs) / kTcBufs : 0;
  tc.fu0 += (uint32_t)nrel0;
  tc.fu1 += (uint32_t)nrel1;
  use0 += (uint32_t)((ntl + kTcBufs - 1) / kTcBufs);
  use1 += (uint32_t)(kTcBufs > 1 ? ntl / kTcBufs : 0);
}

#ifdef CB2_GIBBS_PROF
#define GIBBS_T(slot) do { long long t_ = clock64(); gprof[slot] += t_ - glast; glast = t_; } while (0)
#else
#define GIBBS_T(slot) do { } while (0)
#endif

// SMALL (small products: every message of the batch has a few hundred points at most, e.g. the 100 particles of an IIF solve): the
// records of ALL levels of the product's messages are staged in shared memory once, the scoring passes read them in place (no bulk
// copy to wait for, no CTA barrier per tile), the leave-one-out gathers and the second pass read shared memory instead of L2.  A
// (level, sweep, message) phase of a 100-point product costs ~1.7 us otherwise, almost all of it memory latency.  Same arithmetic
// in the same order: bit-identical chains.
template <typename T, int DIM, uint32_t CIRC, bool SMALL = false>
__global__ void __launch_bounds__(kS / ChainsPerThread<T, DIM>::value, SMALL ? 1 : (ChainsPerThread<T, DIM>::value == 2 ? 7 : (sizeof(T) == 4 ? CB2_GIBBS_CTAS : 4)))
gibbs_kernel(const GibbsArgs a) {
  constexpr int RECN = (2 * DIM + 1 + 3) / 4 * 4, RECL = (DIM + 1 + 3) / 4 * 4;
  constexpr int TILE_BYTES = gibbs_tile_bytes<T, DIM>();
  // candidates per tile, rounded down to whole scoring groups: tile_pass closes chunks on multiples of kG and assumes every tile
  // starts on one (d = 4, 5 have 48-byte coarse records: 170 per 8 KB tile would shift every later tile off the grid)
  constexpr int TN_NODE = TILE_BYTES / (RECN * (int)sizeof(T)) / kG * kG;
  constexpr int TN_LEAF = TILE_BYTES / (RECL * (int)sizeof(T)) / kG * kG;
  static_assert(TN_NODE > 0 && TN_LEAF > 0 && TN_NODE % kG == 0 && TN_LEAF % kG == 0, "tile sizes must be whole scoring groups");
  constexpr int R = ChainsPerThread<T, DIM>::value;  // chains per thread
  constexpr int NT = kS / R;                          // threads per CTA; the CTA always advances kS chains
  constexpr bool kPack = DIM == 6 && sizeof(T) == 4;  // packed-pair scoring (see PackedConst)
  const uint32_t circ = CIRC == ~0u ? a.circ_mask : CIRC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* tile[2] = {reinterpret_cast<T*>(smem_raw), reinterpret_cast<T*>(smem_raw + TILE_BYTES)};
  T* chunk = reinterpret_cast<T*>(smem_raw + 2 * TILE_BYTES);                // [kNCH][kS]
  ind_t* ind_s = reinterpret_cast<ind_t*>(chunk + kNCH * kS);                  // [CB2_MAX_DENS][kS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ind_s + CB2_MAX_DENS * kS);    // [2]
  SMeta<T>* sm = reinterpret_cast<SMeta<T>*>(bars + 2);                        // [CB2_MAX_DENS]
  constexpr bool kTC = CB2_GIBBS_TC && DIM == 6 && sizeof(T) == 4 && R == 1;   // tensor-core leaf passes (leaf_pass_tc)
  TcState tc;
  uint32_t* tmem_slot = nullptr;
  if constexpr (kTC) {
    unsigned char* after = reinterpret_cast<unsigned char*>(sm + CB2_MAX_DENS);
    after = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(after) + 127) & ~(uintptr_t)127);
    tc.sA = reinterpret_cast<float*>(after);
    tc.mbars = reinterpret_cast<uint64_t*>(after + 16384);
    tmem_slot = reinterpret_cast<uint32_t*>(after + 16384 + 32);
    tc.mu0 = tc.mu1 = tc.fu0 = tc.fu1 = 0;
  }

  const int tid = threadIdx.x;
  const Work wk = a.work[blockIdx.x];
  const ProdDev* __restrict__ prp = &a.prods[wk.prod];
  const int D = prp->D, L = prp->L, sweeps = prp->sweeps, Nout = prp->Nout;
  const uint32_t pstream = prp->stream;
  int col[R];        // column of chain r in the per-chain shared arrays
  bool active[R];
  uint32_t sid[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    col[r] = r * NT + tid;
    active[r] = wk.s0 + col[r] < Nout;
    sid[r] = (uint32_t)(prp->sample_offset + wk.s0 + col[r]);
  }
  const T* pool = reinterpret_cast<const T*>(a.node_pool);
  const T HL = (T)(0.5 * CB2_LOG2E);
  const bool single = D == 1;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    if constexpr (kTC) {
      mbar_init(&tc.mbars[0], 1); mbar_init(&tc.mbars[1], 1);
      mbar_init(&tc.mbars[2], NT / 32); mbar_init(&tc.mbars[3], NT / 32);  // accumulator-free: one arrival per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (kTC) {
    if (tid < 32) {  // 128 accumulator columns (two 64-leaf buffers); four CTAs share the SM's 512
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (tid < D) {
    const DensMeta& dm = a.metas[prp->dens[tid]];
    SMeta<T> m;
    for (int l = 0; l <= kMaxLev; ++l) m.rec_off32[l] = (uint32_t)(dm.rec_off[l] >> 5);
    m.feat_off32 = dm.feat_off >= 0 ? (uint32_t)(dm.feat_off >> 5) : 0xffffffffu;
    m.perm_off = dm.perm_off;
    for (int i = 0; i < CB2_MAX_DIM; ++i) m.bw2[i] = (T)(dm.bw[i] * dm.bw[i]);
    m.N = dm.N; m.depth = dm.depth; m.uniform = dm.w == nullptr;
    {
      const double* rg = a.stat + (size_t)prp->dens[tid] * a.stat_stride + (a.stat_stride - 16);
      for (int i = 0; i < CB2_MAX_DIM; ++i) { m.cmin[i] = (float)rg[i] - 1e-3f; m.cmax[i] = (float)rg[8 + i] + 1e-3f; }
    }
    sm[tid] = m;
  }
  for (int k = 0; k < D; ++k)
#pragma unroll
    for (int r = 0; r < R; ++r) ind_s[k * kS + col[r]] = 0;
  __syncthreads();
  if constexpr (kTC) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tc.tmem = *tmem_slot;
  }
  // SMALL: every level of every message into shared memory (behind the regular arrays; launch_gibbs_inst sizes the allocation)
  const T* srec = nullptr;
  if constexpr (SMALL) {
    T* stage = reinterpret_cast<T*>(smem_raw + ((gibbs_smem_bytes_c<T, DIM>() + 127) & ~(size_t)127));
    auto block_elems = [&](const SMeta<T>& m) {  // levels are laid out back to back, each padded to 32 elements (validate_and_plan)
      return (uint32_t)(m.rec_off(m.depth) - m.rec_off(0)) + (uint32_t)((m.N * RECL + 31) & ~31);
    };
    if (tid == 0) {
      uint32_t o = 0;
      for (int k = 0; k < D; ++k) { sm[k].srec32 = o >> 5; o += block_elems(sm[k]); }
    }
    __syncthreads();
    constexpr int V = 16 / (int)sizeof(T);
    for (int k = 0; k < D; ++k) {
      const uint4* src = reinterpret_cast<const uint4*>(pool + sm[k].rec_off(0));
      uint4* dst = reinterpret_cast<uint4*>(stage + (size_t)sm[k].srec32 * 32);
      const int nv = (int)(block_elems(sm[k]) / V);
      for (int e = tid; e < nv; e += NT) dst[e] = __ldg(src + e);
    }
    __syncthreads();
    srec = stage;
  }
  // records of level `lev` of message k: staged copy or node pool
  auto level_rec = [&](int k, int lev) -> const T* {
    if constexpr (SMALL) return srec + (size_t)sm[k].srec32 * 32 + (sm[k].rec_off(lev) - sm[k].rec_off(0));
    else return pool + sm[k].rec_off(lev);
  };
  const T* const cpool = SMALL ? srec : pool;   // what combine() indexes
  uint32_t use0 = 0, use1 = 0;  // completed uses of each tile buffer (phase parity)
#ifdef CB2_GIBBS_PROF
  long long gprof[6] = {0, 0, 0, 0, 0, 0}, glast = clock64();
#endif

  for (int step = 1; step <= L; ++step) {
    // levelDown: a leaf stays itself, otherwise move to one of the node's children, drawn in proportion to their weights (one
    // uniform per chain, message and level: Philox tag 2 -- the (Niter + 2)-th uniform upstream reserves per level, see
    // oracle/cb2_oracle.c).  Always descending to the first child would pin every chain to the left half of a tree whose halves
    // do not overlap.
    for (int k = 0; k < D; ++k) {
      const int depth = sm[k].depth;
      if (step <= depth) {
        const bool toleaf = step == depth;
        const T* crec = level_rec(k, step);
        const int crl = toleaf ? RECL : RECN, cwo = toleaf ? DIM : 2 * DIM;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          int node = ind_s[k * kS + col[r]], c0, nch = 2, lo, n;
          node_range(sm[k].N, step - 1, node, lo, n);  // leaves under the current node
          if (toleaf) { c0 = lo; nch = n; }
          else c0 = 2 * node;
          int pick = c0;
          if (nch > 1 && a.level_down) {
            T w0, w1;
            if (sm[k].uniform) {  // equal kernel weights: the children weigh their leaf counts (left = ceil(n / 2)), no loads
              w0 = (T)((n + 1) >> 1); w1 = (T)(n - ((n + 1) >> 1));
            } else {
              w0 = fast_exp2(crec[(size_t)c0 * crl + cwo]); w1 = fast_exp2(crec[(size_t)(c0 + 1) * crl + cwo]);
            }
            uint32_t rnd[4];
            philox4x32_10(sid[r], (uint32_t)((step - 1) * D + k), pstream, 2u, a.k0, a.k1, rnd);
            const T u = u01<T>(rnd[0], rnd[1]);
            if (!(w0 >= u * (w0 + w1))) pick = c0 + 1;
          }
          ind_s[k * kS + col[r]] = (ind_t)pick;
        }
      }
    }
    for (int sweep = 0; sweep < sweeps; ++sweep) {
      for (int j = 0; j < D; ++j) {
        const SMeta<T>& dj = sm[j];
        const int lev = min(step, dj.depth);
        const bool leaf = lev == dj.depth;
        const int cnt = level_count(dj.N, dj.depth, lev);
        const bool cw = dj.uniform != 0 && (dj.N & (dj.N - 1)) == 0;  // every node of a coarse level weighs the same
        const T* lrec = level_rec(j, lev);
        const int rec = leaf ? RECL : RECN;
        const int TN = SMALL ? cnt : (leaf ? TN_LEAF : TN_NODE);   // SMALL: the whole level is one "tile", read in place
        const int ntiles = (cnt + TN - 1) / TN;
        int CH = kG;
        while (CH * kNCH < cnt) CH <<= 1;

        GIBBS_T(0);
        ChainConst<T, DIM> cc[R];
        PackedConst pc[R];
        bool nowrap = true;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          T M[DIM], C[DIM];
          if (!single) combine<T, DIM, SMALL>(cpool, sm, D, ind_s, col[r], step, j, circ, M, C);
          // the wrap is a no-op for this chain if every stored coordinate of message j lies within pi of its mean
          if (circ != 0u && !single) {
#pragma unroll
            for (int i = 0; i < DIM; ++i)
              if (circ >> i & 1) nowrap = nowrap && ((float)M[i] - 3.14159f <= dj.cmin[i]) && (dj.cmax[i] <= (float)M[i] + 3.14159f);
          }
          cc[r].qs = single ? T(0) : HL;
          cc[r].cmul = single ? T(0) : T(1);
#pragma unroll
          for (int i = 0; i < DIM; ++i) {
            if (leaf) {
              T h = single ? T(0) : HL / (dj.bw2[i] + C[i]);
              T sh = tsqrt(h);
              cc[r].a[i] = sh; cc[r].b[i] = single ? T(0) : M[i] * sh; cc[r].p[i] = T(CB2_TWO_PI) * sh;
            } else {
              cc[r].a[i] = single ? T(0) : M[i]; cc[r].b[i] = single ? T(1) : C[i]; cc[r].p[i] = T(0);
            }
          }
          pc[r] = make_packed<T, DIM>(cc[r], leaf);
        }

        GIBBS_T(1);
        // warp-uniform choice of the scoring loop (all chains of the warp must agree)
        const bool flat = circ == 0u || __all_sync(0xffffffffu, nowrap);
        bool one_den = false;  // coarse level over one common denominator (score_node_rsq1): every chain of the warp proves the range
        if constexpr (kPack && CB2_COARSE_RSQ && CB2_ONE_DEN && R == 1) {
          if (!leaf && !single) {
            float lo = 1.f, hi = 1.f;
#pragma unroll
            for (int i = 0; i < DIM; ++i) {
              // (circular coordinates: the statistics are taken on the unrotated angles, whose spread is bounded by pi)
              const float half = (circ >> i & 1) ? 3.1416f : 0.5f * (dj.cmax[i] - dj.cmin[i]);
              lo *= (float)dj.bw2[i] + (float)cc[0].b[i];
              hi *= fmaf(half, half, (float)dj.bw2[i]) + (float)cc[0].b[i];
            }
            one_den = __all_sync(0xffffffffu, lo > 1e-30f && hi < 1e30f);
          }
        }
        // ---- pass 1: chunk partial sums of exp2(e - m), m = running (lazily raised) reference maximum
        T m[R], csum[R], thr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { m[r] = T(0); csum[r] = T(0); thr[r] = T(60); }  // m is set from the level's first candidate (tile_pass / leaf_pass_tc)
        int cidx = 0;
        bool done_tc = false;
        if constexpr (kTC) {
          if (leaf && dj.feat_off32 != 0xffffffffu) {  // CTA-uniform conditions: every thread reaches the vote
            const bool cta_flat = circ == 0u || __syncthreads_and(nowrap ? 1 : 0) != 0;
            if (cta_flat) {
              if (CH < 32) CH = 32;  // the tensor-core pass sums blocks of 32 leaves (pass 2 below scans the same chunks)
              leaf_pass_tc<T, DIM>(reinterpret_cast<const float*>(pool) + (size_t)dj.feat_off32 * 32, cnt, CH, cc[0], dj.uniform != 0, smem_raw, TILE_BYTES,
                                   bars, use0, use1, tc, reinterpret_cast<float*>(chunk), tid, m[0], csum[0], cidx,
                                   reinterpret_cast<const float*>(lrec) + (size_t)ind_s[j * kS + tid] * RECL);
              done_tc = true;
            }
          }
        }
        if (!SMALL && !done_tc && tid == 0) {
          uint32_t bytes = (uint32_t)(min(TN, cnt) * rec * (int)sizeof(T));
          mbar_expect_tx(&bars[0], bytes);
          tma_load_1d(tile[0], lrec, bytes, &bars[0]);
        }
        for (int t = 0; t < (done_tc ? 0 : ntiles); ++t) {
          const int b = t & 1;
          if constexpr (!SMALL) {
            if (tid == 0 && t + 1 < ntiles) {
              int n1 = min(TN, cnt - (t + 1) * TN);
              uint32_t bytes = (uint32_t)(n1 * rec * (int)sizeof(T));
              mbar_expect_tx(&bars[b ^ 1], bytes);
              tma_load_1d(tile[b ^ 1], lrec + (size_t)(t + 1) * TN * rec, bytes, &bars[b ^ 1]);
            }
            mbar_wait(&bars[b], (b ? use1 : use0) & 1);
            if (b) ++use1; else ++use0;
          }
          const int nt = min(TN, cnt - t * TN);
          const T* tb = SMALL ? lrec : reinterpret_cast<const T*>(smem_raw + b * TILE_BYTES);
          // with the tensor-core leaf pass compiled in, the packed leaf loops would only serve warp-flat passes of CTAs that are
          // not flat as a whole: they take the wrapped loops instead and the kernel stays smaller (instruction cache)
          if (flat && !(kTC && leaf)) {  // Euclidean arithmetic on every coordinate: same values as the wrapped form, fewer instructions
            if (!leaf) {
              if constexpr (kPack && CB2_COARSE_RSQ && CB2_ONE_DEN && R == 1) {
                if (one_den) {
                  if (cw) tile_pass<T, DIM, 5, kPack, R>(tb, nt, t * TN, CH, cc, pc, 0u, chunk, tid, m, csum, cidx, thr);
                  else tile_pass<T, DIM, 4, kPack, R>(tb, nt, t * TN, CH, cc, pc, 0u, chunk, tid, m, csum, cidx, thr);
                } else if (cw) tile_pass<T, DIM, 3, kPack, R>(tb, nt, t * TN, CH, cc, pc, 0u, chunk, tid, m, csum, cidx, thr);
                else tile_pass<T, DIM, 0, kPack, R>(tb, nt, t * TN, CH, cc, pc, 0u, chunk, tid, m, csum, cidx, thr);
              } else {
                if (kPack && cw) tile_pass<T, DIM, 3, kPack, R>(tb, nt, t * TN, CH, cc, pc, 0u, chunk, tid, m, csum, cidx, thr);
                else tile_pass<T, DIM, 0, kPack, R>(tb, nt, t * TN, CH, cc, pc, 0u, chunk, tid, m, csum, cidx, thr);
              }
            }
            else if constexpr (!kTC) {
              if (!dj.uniform) tile_pass<T, DIM, 1, kPack, R>(tb, nt, t * TN, CH, cc, pc, 0u, chunk, tid, m, csum, cidx, thr);
              else tile_pass<T, DIM, 2, kPack, R>(tb, nt, t * TN, CH, cc, pc, 0u, chunk, tid, m, csum, cidx, thr);
            }
          } else {
            if (!leaf) tile_pass<T, DIM, 0, false, R>(tb, nt, t * TN, CH, cc, pc, circ, chunk, tid, m, csum, cidx, thr);
            else if (!dj.uniform) tile_pass<T, DIM, 1, false, R>(tb, nt, t * TN, CH, cc, pc, circ, chunk, tid, m, csum, cidx, thr);
            else tile_pass<T, DIM, 2, false, R>(tb, nt, t * TN, CH, cc, pc, circ, chunk, tid, m, csum, cidx, thr);
          }
          if constexpr (!SMALL) __syncthreads();  // everyone is done with tile[b] before it is refilled
        }
        if (cnt & (CH - 1)) {
#pragma unroll
          for (int r = 0; r < R; ++r) chunk[cidx * kS + col[r]] = csum[r];
          ++cidx;
        }
        const int nch = cidx;
        GIBBS_T(leaf ? 3 : 2);

        // ---- inverse CDF: pick the chunk, then re-score only that chunk
#pragma unroll
        for (int r = 0; r < R; ++r) {
          T tot = T(0);
          for (int c = 0; c < nch; ++c) tot += chunk[c * kS + col[r]];
          uint32_t rnd[4];
          const uint32_t draw = (uint32_t)(((step - 1) * sweeps + sweep) * D + j);
          philox4x32_10(sid[r], draw, pstream, 0u, a.k0, a.k1, rnd);
          const T tgt = u01<T>(rnd[0], rnd[1]) * tot;
          T cum = T(0);
          int sel = nch - 1;
          for (int c = 0; c < nch; ++c) {
            T v = chunk[c * kS + col[r]];
            if (cum + v >= tgt) { sel = c; break; }
            cum += v;
          }
          const int z0 = sel * CH, z1 = min(cnt, z0 + CH);
          GIBBS_T(4);
          // pass 2 must reproduce the scores of pass 1 bit for bit, so it takes the same (packed or wrapped) scorer
          int pick;
          if (flat && !(kTC && leaf)) {
            if (!leaf) {
              if constexpr (kPack && CB2_COARSE_RSQ && CB2_ONE_DEN && R == 1) {
                if (one_den) pick = cw ? chunk_pick<T, DIM, 5, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt)
                                       : chunk_pick<T, DIM, 4, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt);
                else pick = cw ? chunk_pick<T, DIM, 3, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt)
                               : chunk_pick<T, DIM, 0, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt);
              } else {
                pick = (kPack && cw) ? chunk_pick<T, DIM, 3, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt)
                                     : chunk_pick<T, DIM, 0, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt);
              }
            }
            else if constexpr (!kTC) {
              pick = !dj.uniform ? chunk_pick<T, DIM, 1, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt)
                                 : chunk_pick<T, DIM, 2, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt);
            } else pick = z0;
          } else if (kTC && done_tc) {  // tensor-core leaf pass: the CTA is wrap-free, the packed FP32 leaf scorer applies
            if constexpr (kTC) {
              pick = !dj.uniform ? chunk_pick<T, DIM, 1, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt)
                                 : chunk_pick<T, DIM, 2, kPack, SMALL>(lrec, z0, z1, cc[r], pc[r], 0u, m[r], cum, tgt);
            } else pick = z0;
          } else {
            pick = !leaf ? chunk_pick<T, DIM, 0, false, SMALL>(lrec, z0, z1, cc[r], pc[r], circ, m[r], cum, tgt)
                         : (!dj.uniform ? chunk_pick<T, DIM, 1, false, SMALL>(lrec, z0, z1, cc[r], pc[r], circ, m[r], cum, tgt)
                                        : chunk_pick<T, DIM, 2, false, SMALL>(lrec, z0, z1, cc[r], pc[r], circ, m[r], cum, tgt));
          }
          if (active[r]) ind_s[j * kS + col[r]] = (ind_t)pick;
          GIBBS_T(5);
        }
      }
    }
  }

#ifdef CB2_GIBBS_PROF
  if (blockIdx.x == gridDim.x / 2 && (tid & 31) == 0)  // a CTA from the middle of the grid: the SM is fully loaded
    printf("gibbs prof warp %d (clk): other %lld combine %lld pass1-coarse %lld pass1-leaf %lld cdf %lld pass2 %lld\n", tid >> 5, gprof[0], gprof[1], gprof[2], gprof[3], gprof[4], gprof[5]);
#endif
  // final point: product of all selected kernels plus one Gaussian draw; coordinates were stored relative to the centre of
  // the product's first message (tree_build_kernel: Euclidean always, circular in FP32 mode)
  const double* cen0 = a.metas[prp->dens[0]].center;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (!active[r]) continue;
    const int s = wk.s0 + col[r];
    T M[DIM], C[DIM];
    combine<T, DIM, SMALL>(cpool, sm, D, ind_s, col[r], L, -1, circ, M, C);
    uint32_t rnd[4];
    double* out = prp->out;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      if ((i & 1) == 0) philox4x32_10(sid[r], (uint32_t)(i >> 1), pstream, 1u, a.k0, a.k1, rnd);
      double u1 = (double)u01<T>(rnd[0], rnd[1]), u2 = (double)u01<T>(rnd[2], rnd[3]);
      double rr = sqrt(-2.0 * log(u1)), sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      double v = (double)M[i] + (((circ >> i & 1) && sizeof(T) == 8) ? 0.0 : cen0[i]) + sqrt((double)C[i]) * ((i & 1) ? rr * sn : rr * cs);
      if (circ >> i & 1) v = wrap_pi(v);
      out[(size_t)s * DIM + i] = v;
    }
    int* labels = prp->labels;
    if (labels) {
      for (int k = 0; k < D; ++k) labels[(size_t)s * D + k] = a.int_pool[sm[k].perm_off + ind_s[k * kS + col[r]]];
    }
  }
  if constexpr (kTC) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tc.tmem), "r"(128u));
  }
}

template <typename T, int DIM> size_t gibbs_smem_bytes() { return gibbs_smem_bytes_c<T, DIM>(); }
