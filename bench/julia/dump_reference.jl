# dump_reference.jl -- run the REAL upstream implementation of the hot path on fixed inputs and write inputs + outputs to
# tests/golden/julia_reference.json, so that the CPU oracle (oracle/) can be pinned against Julia instead of against SURVEY App. A.
#
# The build container of this repository has no Julia, and the arithmetic lives in packages that are not in the Caesar.jl tree
# (ApproxManifoldProducts, KernelDensityEstimate, IncrementalInference, Manifolds, RoME).  Run this once on any machine with
#   julia> ] add Caesar RoME IncrementalInference ApproxManifoldProducts KernelDensityEstimate Manifolds JSON3 StaticArrays
#   $ JULIA_NUM_THREADS=auto julia bench/julia/dump_reference.jl [path/to/Caesar.jl] [path/to/this/repo]
# and commit the JSON; tests/test_oracle_julia_golden.py activates when the file exists.
#
# Every section stores the exact inputs it used (no RNG has to match across languages) and is wrapped in try/catch, so one API
# drift does not lose the other sections.  Reference call sites reproduced:
#   mmd(M, a, b, N, M, threads; bw)          Caesar ext/Images/ScatterAlignPose2.jl:179
#   mmd(mkdA, mkdB)                          Caesar test/ICRA2022_tests/insufficient_data.jl:47-57
#   manikde!(M, pts)  (LOO bandwidth)        Caesar test/testScatterAlignPose2.jl:187
#   (mkd)(pts)                               Caesar test/ICRA2022_tests/multihypo_corners.jl:66-90
#   manifoldProduct(ff, M; Niter, N)         Caesar ext/services/additional.jl:20-28
#   sample(mkd, n)                           Caesar ext/Images/ScatterAlignPose2.jl:140-141
#   _PCL._transformPointCloud(M, pts, c; backward)   Caesar src/3rdParty/_PCL/services/ConsolidateRigidTransform.jl:17-54
#   factor residual (cf)(X, p, q)            Caesar ext/Images/ScatterAlignPose2.jl:216-231
using Random, Statistics, LinearAlgebra
using JSON3, StaticArrays
using Manifolds
using ApproxManifoldProducts
import ApproxManifoldProducts as AMP
using KernelDensityEstimate
using IncrementalInference, RoME
using Caesar

caesar_dir = length(ARGS) >= 1 ? ARGS[1] : pkgdir(Caesar)
repo_dir = length(ARGS) >= 2 ? ARGS[2] : normpath(joinpath(@__DIR__, "..", ".."))
out = Dict{String,Any}("julia_version" => string(VERSION), "threads" => Threads.nthreads())
for (name, mod) in (("Caesar", Caesar), ("ApproxManifoldProducts", AMP), ("KernelDensityEstimate", KernelDensityEstimate),
                    ("IncrementalInference", IncrementalInference), ("Manifolds", Manifolds), ("RoME", RoME))
  out["version_" * name] = try string(pkgversion(mod)) catch; "unknown" end
end
rows(v) = [collect(Float64, x) for x in v]                       # Vector of points -> JSON rows
mat(A) = [collect(Float64, A[:, j]) for j in 1:size(A, 2)]        # d x N matrix -> N rows of d
section(f, name) = try out[name] = f(); println("ok    ", name) catch err; out[name * "_error"] = sprint(showerror, err); println("FAILED ", name, ": ", err) end

Random.seed!(20240001)

# ---- fixtures: the 44 stored beliefs of test/ICRA2022_tests/test_data/tut3
fixdir = joinpath(caesar_dir, "test", "ICRA2022_tests", "test_data", "tut3")
loadbel(n) = unpackDistribution(JSON3.read(read(joinpath(fixdir, n * ".json"), String), IncrementalInference.PackedManifoldKernelDensity))
pairs_ = [("l1_1", "l1_2"), ("l3_2", "l3_3"), ("x1_2", "x1_3"), ("l1_1", "l2_1"), ("x0_1", "x0_2"), ("l4_4", "l4_5")]

# ---- 1. mmd between stored beliefs: mmd(mkdA, mkdB) with its default bw, and the positional form on raw points
section("mmd_fixture_pairs") do
  res = []
  for (a, b) in pairs_
    A, B = loadbel(a), loadbel(b)
    pa, pb = getPoints(A), getPoints(B)
    v_default = mmd(A, B)
    M = TranslationGroup(2)
    v_pos = [mmd(M, pa, pb, length(pa), length(pb), false; bw = SA[bw]) for bw in (0.001, 0.05, 1.0)]
    push!(res, Dict("a" => a, "b" => b, "mmd_default" => v_default, "bw" => [0.001, 0.05, 1.0], "mmd_positional" => v_pos))
  end
  res
end

# ---- 2. mmd on SE(2) beliefs (the metric question of SURVEY App. A.1 item 4): points given as coordinates (x, y, theta)
section("mmd_se2") do
  M = SpecialEuclidean(2)
  e0 = ArrayPartition(SA[0.0, 0.0], SMatrix{2,2}(1.0, 0.0, 0.0, 1.0))
  ca = [[randn() * 2, randn(), 0.4 * randn()] for _ in 1:60]
  cb = [[randn() * 2 + 0.5, randn(), 0.4 * randn() + 0.3] for _ in 1:50]
  pt(c) = exp(M, e0, hat(M, e0, c))
  pa, pb = pt.(ca), pt.(cb)
  vals = [mmd(M, pa, pb, length(pa), length(pb), false; bw = SA[bw]) for bw in (0.001, 0.1, 1.0)]
  A = manikde!(M, pa); B = manikde!(M, pb)
  Dict("coords_a" => ca, "coords_b" => cb, "bw" => [0.001, 0.1, 1.0], "mmd" => vals, "mmd_mkd_default" => mmd(A, B),
       "bw_a" => collect(getBW(A)[:, 1]), "bw_b" => collect(getBW(B)[:, 1]))
end

# ---- 3. manikde! bandwidths (LOO cross-validation) on fresh inputs: Point2, ContinuousScalar, Pose2
section("manikde_bandwidth") do
  res = []
  for (name, M, gen) in (("Point2_bimodal", TranslationGroup(2), () -> [SA[randn() + 4 * (rand() < 0.5), 0.3 * randn()] for _ in 1:150]),
                         ("Scalar_narrow", TranslationGroup(1), () -> [SA[0.05 * randn()] for _ in 1:80]),
                         ("Point3", TranslationGroup(3), () -> [SA[randn(), 2 * randn(), 0.1 * randn()] for _ in 1:200]))
    pts = gen()
    b = manikde!(M, pts)
    push!(res, Dict("name" => name, "pts" => rows(pts), "bw" => collect(getBW(b)[:, 1])))
  end
  res
end

# ---- 4. evaluate a belief at query points: (mkd)(pts)
section("kde_evaluate") do
  res = []
  for n in ("l1_1", "x0_1", "l3_3")
    B = loadbel(n)
    q = [SA[randn() * 20, randn() * 20] for _ in 1:40]
    append!(q, getPoints(B)[1:10])
    Q = hcat(collect.(q)...)                       # d x Q matrix as the reference tests pass it
    push!(res, Dict("belief" => n, "queries" => rows(q), "density" => collect(Float64, B(Q))))
  end
  res
end

# ---- 5. manifoldProduct: sample moments over repeated draws (the product is stochastic; the oracle must match in distribution)
function product_moments(M, ff; N, reps)
  allpts = Vector{Vector{Float64}}()
  for _ in 1:reps
    p = manifoldProduct(ff, M; Niter = 1, N = N)
    append!(allpts, [collect(Float64, AMP.makeCoordsFromPoint(M, x)) for x in getPoints(p)])
  end
  X = hcat(allpts...)
  Dict("N" => N, "reps" => reps, "mean" => vec(mean(X, dims = 2)), "cov" => mat(cov(X')), "samples_first" => allpts[1:min(end, 400)])
end
section("manifold_product") do
  res = []
  # (a) two stored Point2 beliefs; (b) three 1-D Gaussians (closed form known); (c) the ZMQ example of ext/services/additional.jl:34
  A, B = loadbel("l1_1"), loadbel("l1_2")
  push!(res, merge(Dict("name" => "tut3_l1_1_x_l1_2", "inputs" => [Dict("pts" => rows(getPoints(A)), "bw" => collect(getBW(A)[:, 1])),
                                                                  Dict("pts" => rows(getPoints(B)), "bw" => collect(getBW(B)[:, 1]))]),
                   product_moments(TranslationGroup(2), [A, B]; N = 100, reps = 200)))
  M1 = TranslationGroup(1)
  gs = [manikde!(M1, [SA[m + s * randn()] for _ in 1:200]) for (m, s) in ((0.0, 1.0), (1.0, 0.5), (-0.5, 2.0))]
  push!(res, merge(Dict("name" => "three_gaussians_1d", "inputs" => [Dict("pts" => rows(getPoints(g)), "bw" => collect(getBW(g)[:, 1])) for g in gs]),
                   product_moments(M1, gs; N = 200, reps = 200)))
  # bimodal x unimodal on Pose2 (wrapped angle near +-pi)
  M = SpecialEuclidean(2); e0 = ArrayPartition(SA[0.0, 0.0], SMatrix{2,2}(1.0, 0.0, 0.0, 1.0)); pt(c) = exp(M, e0, hat(M, e0, c))
  c1 = [[randn() + 3 * (rand() < 0.5), 0.5 * randn(), 3.0 + 0.2 * randn()] for _ in 1:150]
  c2 = [[0.8 * randn() + 3, 0.5 * randn(), -3.0 + 0.3 * randn()] for _ in 1:150]
  P1, P2 = manikde!(M, pt.(c1)), manikde!(M, pt.(c2))
  push!(res, merge(Dict("name" => "pose2_wrapped", "inputs" => [Dict("coords" => c1, "bw" => collect(getBW(P1)[:, 1])), Dict("coords" => c2, "bw" => collect(getBW(P2)[:, 1]))]),
                   product_moments(M, [P1, P2]; N = 150, reps = 200)))
  res
end

# ---- 6. sample(mkd, n): moments of fresh samples
section("sample") do
  B = loadbel("x1_3")
  s, = sample(B, 20000)
  X = hcat(collect.(s)...)
  Dict("belief" => "x1_3", "n" => 20000, "mean" => vec(mean(X, dims = 2)), "cov" => mat(cov(X')))
end

# ---- 7. rigid transform of a cloud and the ScatterAlign cost at fixed coordinates (no optimiser)
section("transform_and_cost") do
  res = []
  for (M, d, nc) in ((SpecialEuclidean(2), 2, 3), (SpecialEuclidean(3), 3, 6))
    pts = [SVector{d}(randn(d)) for _ in 1:30]
    ref = [SVector{d}(randn(d)) for _ in 1:25]
    for c in ([2.0; 0.0; zeros(nc - 3); pi / 6][1:nc], 0.3 .* randn(nc))
      for backward in (false, true)
        tp = Caesar._PCL._transformPointCloud(M, pts, c; backward = backward)
        cost = mmd(M.manifold[1], ref, tp, length(ref), length(tp), false; bw = SA[1.0])
        push!(res, Dict("dim" => d, "coords" => collect(c), "backward" => backward, "pts" => rows(pts), "ref" => rows(ref),
                        "transformed" => rows(tp), "mmd_bw1" => cost))
      end
    end
  end
  res
end

# ---- 8. factor residual of ScatterAlignPose2 (the Pose2Pose2 body) at fixed (X, p, q), and the exp / log conventions it rests on
section("residual_pose2") do
  M = SpecialEuclidean(2); e0 = ArrayPartition(SA[0.0, 0.0], SMatrix{2,2}(1.0, 0.0, 0.0, 1.0))
  res = []
  for _ in 1:12
    Xc, pc, qc = [randn(), randn(), 0.5 * randn()], [3 * randn(), 3 * randn(), 2 * randn()], [3 * randn(), 3 * randn(), 2 * randn()]
    X = hat(M, e0, Xc); p = exp(M, e0, hat(M, e0, pc)); q = exp(M, e0, hat(M, e0, qc))
    qhat = Manifolds.compose(M, p, exp(M, e0, X))
    r = vee(M, q, log(M, q, qhat))
    push!(res, Dict("X" => Xc, "p" => pc, "q" => qc, "residual" => collect(Float64, r),
                    "p_translation" => collect(Float64, submanifold_component(p, 1)), "p_rotation" => mat(Matrix(submanifold_component(p, 2)))))
  end
  res
end

mkpath(joinpath(repo_dir, "tests", "golden"))
open(joinpath(repo_dir, "tests", "golden", "julia_reference.json"), "w") do io
  JSON3.write(io, out)
end
println("wrote ", joinpath(repo_dir, "tests", "golden", "julia_reference.json"))
