# CaesarB200.jl -- the reference-side binding: Julia host code keeps the existing API (`manikde!`, `manifoldProduct`,
# `mmd`, `sample`, evaluation of a ManifoldKernelDensity, the ScatterAlign `getSample`) and reaches the B200 kernels
# through `ccall` into libcaesar_b200.so (C ABI: include/caesar_b200.h).  No CUDA.jl, no multi-backend dispatch, no CPU
# fallback: if the library cannot create a context the methods throw.
#
# NOT EXERCISED IN THE BUILD CONTAINER (Julia is not installed there).  The same entry points, argument order and
# layouts are exercised from Python/ctypes by tests/test_gpu_parity.py; INTEGRATION.md shows where each method hooks
# into ApproxManifoldProducts / IncrementalInference / Caesar.
module CaesarB200

using ApproxManifoldProducts
using ApproxManifoldProducts: ManifoldKernelDensity, getPoints, getBW
using Manifolds, ManifoldsBase
using StaticArrays
import IncrementalInference as IIF

const libcb2 = get(ENV, "CAESAR_B200_LIB", joinpath(@__DIR__, "..", "libcaesar_b200.so"))
const CB2_FP32 = Cint(0)
const CB2_FP64 = Cint(1)

struct CB2Density
  N::Int32
  pts::Ptr{Float64}
  bw::Ptr{Float64}
  w::Ptr{Float64}
end

struct CB2ProductDesc
  dens::Ptr{CB2Density}
  D::Int32
  Nout::Int32
  niter::Int32
  stream::UInt32
  sample_offset::Int32
  pts_out::Ptr{Float64}
  labels_out::Ptr{Int32}
  bw_out::Ptr{Float64}
end

mutable struct Context
  h::Ptr{Cvoid}
  lk::ReentrantLock
end

function Context(; device::Integer = parse(Int, get(ENV, "CAESAR_B200_DEVICE", "0")),
                   precision::Symbol = Symbol(get(ENV, "CAESAR_B200_PRECISION", "fp32")))
  h = Ref{Ptr{Cvoid}}(C_NULL)
  rc = ccall((:cb2_create, libcb2), Cint, (Cint, Cint, Ref{Ptr{Cvoid}}), device, precision == :fp64 ? CB2_FP64 : CB2_FP32, h)
  rc == 0 || error("libcaesar_b200: " * unsafe_string(ccall((:cb2_last_error, libcb2), Cstring, (Ptr{Cvoid},), C_NULL)))
  ctx = Context(h[], ReentrantLock())
  finalizer(c -> ccall((:cb2_destroy, libcb2), Cvoid, (Ptr{Cvoid},), c.h), ctx)
  return ctx
end

const _ctx = Ref{Union{Nothing,Context}}(nothing)
context() = (_ctx[] === nothing && (_ctx[] = Context()); _ctx[]::Context)

# a non-zero status becomes a Julia error, so a clique Task fails the way it does today (SURVEY sec. 5)
function _check(ctx::Context, rc::Cint)
  rc == 0 && return nothing
  error("libcaesar_b200 error $rc: " * unsafe_string(ccall((:cb2_last_error, libcb2), Cstring, (Ptr{Cvoid},), ctx.h)))
end

# d x N Float64, point-contiguous: a Matrix{Float64} as is; Vector{<:StaticVector{d,Float64}} is bit-identical
_flat(pts::AbstractMatrix{Float64}) = pts
_flat(pts::AbstractVector{<:StaticVector{D,Float64}}) where {D} = reinterpret(reshape, Float64, pts)
_flat(pts::AbstractVector{<:AbstractVector{<:Real}}) = Float64.(reduce(hcat, pts))

# coordinate topology of a manifold / variable type: 1 = circular
_circular(::typeof(TranslationGroup(1))) = UInt8[0]
_circular(::typeof(TranslationGroup(2))) = UInt8[0, 0]
_circular(::typeof(TranslationGroup(3))) = UInt8[0, 0, 0]
_circular(::typeof(SpecialEuclidean(2))) = UInt8[0, 0, 1]
_circular(::typeof(SpecialEuclidean(3))) = UInt8[0, 0, 0, 1, 1, 1]
_circular(M) = zeros(UInt8, manifold_dimension(M))

# ------------------------------------------------------------------------------------------------ mmd
# replaces AMP.mmd(MF, a, b, N, M, threads; bw)  -- call site Caesar ext/Images/ScatterAlignPose2.jl:179
function b200_mmd(a, b; bw::AbstractVector{<:Real} = SA[0.001;], ctx::Context = context())
  A, B = _flat(a), _flat(b)
  out = Ref{Float64}(0.0)
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_mmd, libcb2), Cint,
                      (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Cint, Float64, Ref{Float64}),
                      ctx.h, A, size(A, 2), B, size(B, 2), size(A, 1), Float64(bw[1]), out))
  end
  return out[]
end
ApproxManifoldProducts.mmd(MF::AbstractManifold, a::AbstractVector, b::AbstractVector, N::Int = length(a), M::Int = length(b),
                           threads::Bool = true; bw::AbstractVector{<:Real} = SA[0.001;]) = b200_mmd(a, b; bw)
# beliefs over Pose2 / Pose3 (multihypo_corners.jl:201-202): product-manifold distance, wrapped on the circular coordinates;
# rot_weight = 2 is Manifolds.distance on SpecialEuclidean(2) (d^2 = |dt|^2 + 2 dtheta^2)
function b200_mmd_manifold(M, a, b; bw::AbstractVector{<:Real} = SA[0.001;], rot_weight::Real = 2.0, ctx::Context = context())
  A, B = _flat(a), _flat(b)
  out = Ref{Float64}(0.0)
  circ = _circular(M)
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_mmd_manifold, libcb2), Cint,
                      (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Cint, Ptr{UInt8}, Float64, Float64, Ref{Float64}),
                      ctx.h, A, size(A, 2), B, size(B, 2), size(A, 1), circ, Float64(rot_weight), Float64(bw[1]), out))
  end
  return out[]
end
# coordinates (vee of the log at the identity) of both beliefs; Euclidean variable types keep the flat kernel
function ApproxManifoldProducts.mmd(a::ManifoldKernelDensity, b::ManifoldKernelDensity; bw::AbstractVector{<:Real} = SA[0.001;])
  circ = _circular(a.manifold)
  any(!iszero, circ) || return b200_mmd(getPoints(a.belief), getPoints(b.belief); bw)
  return b200_mmd_manifold(a.manifold, getPoints(a.belief), getPoints(b.belief); bw)
end

# value + analytic gradient of cost(c) = mmd(a, T(c)^-1 b) for K coordinate vectors (columns of `coords`);
# lets getSample call Optim.optimize(f, g!, x0, BFGS()) instead of finite differences (ScatterAlignPose2.jl:182)
function b200_mmd_se_batch(a, b, coords::AbstractMatrix{Float64}; bw::Real, backward::Bool = true, ctx::Context = context())
  A, B = _flat(a), _flat(b)
  K = size(coords, 2)
  vals = zeros(K); grads = zeros(size(coords))
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_mmd_se_batch, libcb2), Cint,
                      (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Cint, Float64, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
                      ctx.h, A, size(A, 2), B, size(B, 2), size(A, 1), Float64(bw), coords, K, backward, vals, grads))
  end
  return vals, grads
end

# ------------------------------------------------------------------------------------------------ manikde! / evaluate / sample
# replaces the LOO-CV bandwidth search inside AMP.manikde!(M, pts) when `bw` is not given
function b200_bandwidth(M, pts; ctx::Context = context())
  P = _flat(pts)
  d, N = size(P)
  sigma = zeros(d)
  circ = _circular(M)
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_kde_bw_loo, libcb2), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Ptr{UInt8}, Ptr{Float64}),
                      ctx.h, P, d, N, circ, sigma))
  end
  return sigma
end

# (mkd)(pts): flat-coordinate evaluation, exact sum (upstream: dual-tree approximation, rel. tol 1e-3)
function b200_evaluate(mkd::ManifoldKernelDensity, q::AbstractMatrix{Float64}; logp::Bool = false, ctx::Context = context())
  mu = getPoints(mkd.belief); sigma = vec(getBW(mkd.belief)[:, 1])
  out = zeros(size(q, 2))
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_kde_eval, libcb2), Cint,
                      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}),
                      ctx.h, mu, sigma, C_NULL, size(mu, 1), size(mu, 2), q, size(q, 2), logp, out))
  end
  return out
end

function b200_sample(mkd::ManifoldKernelDensity, n::Integer; seed::UInt64 = rand(UInt64), stream::Integer = 0, ctx::Context = context())
  mu = getPoints(mkd.belief); sigma = vec(getBW(mkd.belief)[:, 1])
  d, N = size(mu)
  out = zeros(d, n)
  circ = _circular(mkd.manifold)
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_kde_sample, libcb2), Cint,
                      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint, Ptr{UInt8}, Cint, UInt64, UInt32, Ptr{Float64}),
                      ctx.h, mu, sigma, C_NULL, d, N, circ, n, seed, UInt32(stream), out))
  end
  return out   # coordinates; the caller maps them to manifold points exactly as AMP.sample does
end

# ------------------------------------------------------------------------------------------------ products
# One batched pass over B independent products (all variables of a Gibbs sweep / all ready cliques).
# groups[b] is the Vector{ManifoldKernelDensity} that IIF.propagateBelief hands to AMP.manifoldProduct.
function b200_products(groups::Vector{<:Vector{<:ManifoldKernelDensity}}, M; N::Vector{Int}, Niter::Int = 1,
                       seed::UInt64 = rand(UInt64), streams = collect(0:length(groups)-1), ctx::Context = context())
  B = length(groups)
  d = manifold_dimension(M)
  circ = _circular(M)
  keep = Any[]
  dens = [Vector{CB2Density}(undef, length(g)) for g in groups]
  outs = [zeros(d, N[b]) for b in 1:B]
  labels = [zeros(Int32, length(groups[b]), N[b]) for b in 1:B]
  bws = [zeros(d) for _ in 1:B]
  descs = Vector{CB2ProductDesc}(undef, B)
  for b in 1:B
    for (j, mkd) in enumerate(groups[b])
      pts = getPoints(mkd.belief); bw = vec(getBW(mkd.belief)[:, 1])
      push!(keep, pts, bw)
      dens[b][j] = CB2Density(size(pts, 2), pointer(pts), pointer(bw), C_NULL)
    end
    descs[b] = CB2ProductDesc(pointer(dens[b]), length(groups[b]), N[b], Niter, UInt32(streams[b]), 0,
                              pointer(outs[b]), pointer(labels[b]), pointer(bws[b]))
  end
  GC.@preserve keep dens outs labels bws descs begin
    lock(ctx.lk) do
      _check(ctx, ccall((:cb2_product_batch, libcb2), Cint, (Ptr{Cvoid}, Ptr{CB2ProductDesc}, Cint, Ptr{UInt8}, Cint, UInt64),
                        ctx.h, descs, B, circ, d, seed))
    end
  end
  return outs, bws, labels
end

# drop-in for AMP.manifoldProduct(ff, M; Niter, N): Gibbs product on the device, fresh LOO bandwidth, new MKD
function ApproxManifoldProducts.manifoldProduct(ff::AbstractVector{<:ManifoldKernelDensity}, M::AbstractManifold = ff[1].manifold;
                                                Niter::Int = 1, N::Int = maximum(Npts.(ff)), kw...)
  length(ff) == 1 && return ff[1]                     # upstream early-out: one input is returned unchanged
  if any(f -> f._partial !== nothing, ff)             # partial-dimension inputs: dimension-wise products, as upstream does
    return _partial_product(ff, M; Niter, N)
  end
  outs, bws, _ = b200_products([collect(ff)], M; N = [N], Niter)
  e0 = identity_element(M)
  pts = [exp(M, e0, hat(M, e0, view(outs[1], :, i))) for i in 1:N]
  return manikde!(M, pts; bw = bws[1])
end

# Inputs that inform only some coordinates (`_partial`): coordinate i of the result is the 1-D product of the marginals of every
# input that informs it; a coordinate informed by a single input is a fresh draw of N samples from that marginal (device
# sampler, Philox), as the Python mirror does (caesar.jl_b200/__init__.py `_manifoldProductPartial`).
function _partial_product(ff, M; Niter::Int, N::Int, ctx::Context = context())
  d = manifold_dimension(M)
  circ = _circular(M)
  cols = Vector{Vector{Float64}}(undef, d); bws = zeros(d)
  for i in 1:d
    use = [f for f in ff if f._partial === nothing || i in f._partial]
    isempty(use) && (use = [ff[1]])
    M1 = circ[i] != 0 ? RealCircleGroup() : TranslationGroup(1)
    marg = [manikde!(M1, [SA[c[i]] for c in eachcol(_coords(getPoints(f.belief)))]; bw = [getBW(f.belief)[i, 1]]) for f in use]
    if length(marg) == 1
      cols[i] = vec(b200_sample(marg[1], N; ctx)); bws[i] = getBW(marg[1].belief)[1, 1]
    else
      o, b, _ = b200_products([marg], M1; N = [N], Niter, ctx)
      cols[i] = vec(o[1]); bws[i] = b[1][1]
    end
  end
  e0 = identity_element(M)
  pts = [exp(M, e0, hat(M, e0, [cols[i][k] for i in 1:d])) for k in 1:N]
  return manikde!(M, pts; bw = bws)
end

# ------------------------------------------------------------------------------------------------ approxConv (closed form)
# Proposals of closed-form relative factors (kinds 0 PriorPose2, 1 Pose2Pose2, 2 Pose2Point2BearingRange, 3 PriorPose3, 4 Pose3Pose3,
# 5 PriorPoint2, 6 Point2Point2Range, 7 Prior / LinearRelative with caller-sampled noise in `aux`), all factors of a Gibbs colour in one launch; hook: IIF.approxConvBelief / sampleFactor for
# these factor types.
struct CB2ConvDesc
  kind::Int32
  reverse::Int32
  N::Int32
  n_src::Int32
  src::Ptr{Float64}
  aux::Ptr{Float64}
  n_aux::Int32
  stream::UInt32
  mean::NTuple{6,Float64}      # Pose2 / Point2 kinds use the first 3 entries (3 x 3 layout), Pose3 kinds (3, 4) all 6
  chol::NTuple{36,Float64}     # lower Cholesky factor, row-major n x n in the first n^2 entries (n = 3 | 6)
  out::Ptr{Float64}
end
_pad(v, n) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, n)
CB2ConvDesc(kind, reverse, N, src, n_src, aux, n_aux, stream, mean, chol, out) =
  CB2ConvDesc(Int32(kind), Int32(reverse), Int32(N), Int32(n_src), src, aux, Int32(n_aux), UInt32(stream), _pad(mean, 6),
              _pad(vec(permutedims(chol)), 36), out)   # Julia matrices are column-major: the library reads L row-major

function b200_approx_conv!(descs::Vector{CB2ConvDesc}; seed::UInt64 = rand(UInt64), ctx::Context = context())
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_approx_conv_batch, libcb2), Cint, (Ptr{Cvoid}, Ptr{CB2ConvDesc}, Cint, UInt64),
                      ctx.h, descs, length(descs), seed))
  end
  return nothing
end

# ------------------------------------------------------------------------------------------------ factor residual
# (cf::CalcFactor{<:ScatterAlignPose2/3})(X, p, q) (ext/Images/ScatterAlignPose2.jl:216-231, the Pose2Pose2 body) for K particles in
# one launch.  X, p, q: ncoord x K coordinate matrices ([x, y, theta] or [t; rotation vector]).  A single evaluation inside a
# per-particle nonlinear solve should stay the three lines of Julia it is upstream: a kernel launch costs more than the arithmetic.
function b200_se_residual(dim::Int, X::Matrix{Float64}, p::Matrix{Float64}, q::Matrix{Float64}; ctx::Context = context())
  out = similar(X)
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_se_residual_batch, libcb2), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}),
                      ctx.h, dim, X, p, q, size(X, 2), out))
  end
  return out
end

# ------------------------------------------------------------------------------------------------ multi-GPU mmd
# shard `shard` (0-based) of `nshards` of the three pair sums {aa, ab, bb}: the symmetric tile list dealt round-robin.  One
# Julia process per GPU computes its shard; MPI.Allreduce / NCCL over the 3 doubles, then aa/N^2 - 2ab/(NM) + bb/M^2.
function b200_mmd_sums_shard(a, b, shard::Integer, nshards::Integer; bw::AbstractVector{<:Real} = SA[0.001;], ctx::Context = context())
  A, B = _coords(a), _coords(b)
  sums = zeros(3)
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_mmd_sums_shard, libcb2), Cint,
                      (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Cint, Float64, Cint, Cint, Ptr{Float64}),
                      ctx.h, A, size(A, 2), B, size(B, 2), size(A, 1), Float64(bw[1]), shard, nshards, sums))
  end
  return sums
end

# levelDown rule of the multiscale Gibbs sampler (include/caesar_b200.h): 1 = weight-proportional random child (default), 0 = first child
b200_set_level_down(mode::Integer; ctx::Context = context()) =
  lock(() -> _check(ctx, ccall((:cb2_set_level_down, libcb2), Cint, (Ptr{Cvoid}, Cint), ctx.h, mode)), ctx.lk)

# ------------------------------------------------------------------------------------------------ ScatterAlign getSample
# drop-in body for Caesar's getSample(cf::CalcFactor{<:ScatterAlignPose2/3}) (ext/Images/ScatterAlignPose2.jl:155-187):
# K alignments (one per particle of the solver) as one batched BFGS on the device.
function b200_scatter_align(cloud1::ManifoldKernelDensity, cloud2::ManifoldKernelDensity, dim::Int, sample_count::Int, bw::Real,
                            x0::AbstractMatrix{Float64}; seed::UInt64 = rand(UInt64), max_iter::Int = 200, g_tol::Real = 1e-8,
                            ctx::Context = context())
  K = size(x0, 2)
  p1 = getPoints(cloud1.belief); b1 = vec(getBW(cloud1.belief)[:, 1])
  p2 = getPoints(cloud2.belief); b2 = vec(getBW(cloud2.belief)[:, 1])
  coords = zeros(size(x0)); score = zeros(K); iters = zeros(Int32, K)
  GC.@preserve p1 b1 p2 b2 begin
    d1 = Ref(CB2Density(size(p1, 2), pointer(p1), pointer(b1), C_NULL))
    d2 = Ref(CB2Density(size(p2, 2), pointer(p2), pointer(b2), C_NULL))
    lock(ctx.lk) do
      _check(ctx, ccall((:cb2_scatter_align, libcb2), Cint,
                        (Ptr{Cvoid}, Ref{CB2Density}, Ref{CB2Density}, Cint, Cint, Float64, Ptr{Float64}, Cint, UInt64, Cint, Float64,
                         Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                        ctx.h, d1, d2, dim, sample_count, Float64(bw), x0, K, seed, max_iter, Float64(g_tol), coords, score, iters))
    end
  end
  return coords, score, iters
end

# ------------------------------------------------------------------------------------------------ device-resident beliefs
# SURVEY sec. 8f N2: points (d x N) and bandwidths (d) of a belief kept in device buffers between approxConvBelief, manikde!
# and manifoldProduct; every call below is stream-ordered, only b200_download! waits.
struct DeviceBelief
  pts::Ptr{Float64}   # device, d x N
  bw::Ptr{Float64}    # device, d
  N::Int
  d::Int
end

function b200_device_alloc(nbytes::Integer; ctx::Context = context())
  p = Ref{Ptr{Cvoid}}(C_NULL)
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_device_alloc, libcb2), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx.h, nbytes, p))
  end
  return p[]
end
b200_device_free(p::Ptr; ctx::Context = context()) =
  lock(() -> _check(ctx, ccall((:cb2_device_free, libcb2), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, p)), ctx.lk)
b200_upload!(dst::Ptr, src::Array{Float64}; ctx::Context = context()) =
  lock(() -> _check(ctx, ccall((:cb2_upload, libcb2), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Csize_t), ctx.h, dst, src, sizeof(src))), ctx.lk)
b200_download!(dst::Array{Float64}, src::Ptr; ctx::Context = context()) =
  lock(() -> _check(ctx, ccall((:cb2_download, libcb2), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Csize_t), ctx.h, dst, src, sizeof(dst))), ctx.lk)

function DeviceBelief(mkd::ManifoldKernelDensity; ctx::Context = context())
  pts = getPoints(mkd.belief); bw = vec(getBW(mkd.belief)[:, 1])
  dp = convert(Ptr{Float64}, b200_device_alloc(sizeof(pts); ctx)); db = convert(Ptr{Float64}, b200_device_alloc(sizeof(bw); ctx))
  b200_upload!(dp, pts; ctx); b200_upload!(db, bw; ctx)
  return DeviceBelief(dp, db, size(pts, 2), size(pts, 1))
end
function Base.collect(b::DeviceBelief; ctx::Context = context())   # (points, bandwidths) back on the host
  pts = zeros(b.d, b.N); bw = zeros(b.d)
  b200_download!(pts, b.pts; ctx); b200_download!(bw, b.bw; ctx)
  return pts, bw
end

# descriptors carry device pointers (src / aux / out of CB2ConvDesc, every pointer of CB2ProductDesc / CB2Density)
b200_approx_conv_dev!(descs::Vector{CB2ConvDesc}; seed::UInt64 = rand(UInt64), ctx::Context = context()) =
  lock(() -> _check(ctx, ccall((:cb2_approx_conv_batch_dev, libcb2), Cint, (Ptr{Cvoid}, Ptr{CB2ConvDesc}, Cint, UInt64),
                               ctx.h, descs, length(descs), seed)), ctx.lk)
function b200_bandwidth_dev!(sigma_dev::Ptr{Float64}, pts_dev::Vector{Ptr{Float64}}, Ns::Vector{Int32}, M; ctx::Context = context())
  circ = _circular(M)
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_kde_bw_loo_batch_dev, libcb2), Cint,
                      (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Int32}, Cint, Cint, Ptr{UInt8}, Ptr{Float64}),
                      ctx.h, pts_dev, Ns, length(Ns), length(circ), circ, sigma_dev))
  end
end
b200_products_dev!(descs::Vector{CB2ProductDesc}, M; seed::UInt64 = rand(UInt64), ctx::Context = context()) =
  lock(() -> (circ = _circular(M); _check(ctx, ccall((:cb2_product_batch_dev, libcb2), Cint,
                               (Ptr{Cvoid}, Ptr{CB2ProductDesc}, Cint, Ptr{UInt8}, Cint, UInt64),
                               ctx.h, descs, length(descs), circ, length(circ), seed))), ctx.lk)

# ------------------------------------------------------------------------------------------------ ICP branch (sample_count < 0)
# K particles of getSample(cf::CalcFactor{<:ScatterAlignPose3}) (ext/Images/ScatterAlignPose2.jl:187-214) in one call:
# dstb is 6 x K (0.2 * randn(6) per particle, :195); returns the 4 x 4 p_Hicp_phat of every particle and the ICP statistics.
function b200_icp_align(p_pts::AbstractMatrix{Float64}, q_pts::AbstractMatrix{Float64}, dstb::AbstractMatrix{Float64};
                        correspondences::Integer = 500, neighbors::Integer = 50, min_planarity::Real = 0.3, min_change::Real = 3,
                        max_iterations::Integer = 25, ctx::Context = context())
  K = size(dstb, 2)
  Hrm = zeros(4, 4, K); st = zeros(4, K)          # row-major 4 x 4 per particle
  lock(ctx.lk) do
    _check(ctx, ccall((:cb2_icp_align_batch, libcb2), Cint,
                      (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Cint, Cint, Float64, Float64, Cint,
                       Ptr{Float64}, Ptr{Float64}),
                      ctx.h, p_pts, size(p_pts, 2), q_pts, size(q_pts, 2), dstb, K, correspondences, neighbors,
                      Float64(min_planarity), Float64(min_change), max_iterations, Hrm, st))
  end
  return [permutedims(Hrm[:, :, k]) for k in 1:K], st
end

end # module
