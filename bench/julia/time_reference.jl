# time_reference.jl -- CPU baseline of the REAL Julia reference for BASELINE.json's metrics, to be run on the GPU box's host cores
# (the image used for this repository has no Julia; bench.py therefore times the C restatement in oracle/ and says so).
#   $ JULIA_NUM_THREADS=auto julia bench/julia/time_reference.jl
# Prints one JSON line per metric in the shape bench.py's `cpu_baseline` expects ("kind": "reference").
using Random, Statistics, StaticArrays, JSON3, Manifolds
using ApproxManifoldProducts, IncrementalInference, RoME, Caesar
Random.seed!(1000)
nt = Threads.nthreads()

# C5 (SURVEY 8d): D = 8 messages x N = 4096 kernels on Pose3, N_out = 4096, Niter = 1; products are single-threaded upstream, so the
# multithreaded figure is independent variables across Threads.@threads
M = SpecialEuclidean(3)
e0 = ArrayPartition(SA[0.0, 0.0, 0.0], SMatrix{3,3}(1.0I))
function c5_variable()
  truth = [5 .* randn(3); 5 .* (rand(3) .- 0.5)]
  map(1:8) do j
    off = [0.3 .* randn(3); 0.1 .* randn(3)]
    pts = map(1:4096) do _
      sh = (rand() < 0.5 ? 1.0 : -1.0) .* [0.6, 0, 0, 0, 0, 0.2]
      c = truth .+ off .+ sh .+ [0.5, 0.5, 0.5, 0.15, 0.15, 0.15] .* randn(6)
      exp(M, e0, hat(M, e0, c))
    end
    manikde!(M, pts)
  end
end
vars = [c5_variable() for _ in 1:max(nt, 2)]
manifoldProduct(vars[1], M; Niter = 1, N = 64)  # compile
t = @elapsed Threads.@threads for v in vars
  manifoldProduct(v, M; Niter = 1, N = 4096)
end
println(JSON3.write(Dict("metric" => "belief_product_samples_per_s", "value" => length(vars) * 4096 / t, "unit" => "samples/s", "cores" => nt,
                         "kind" => "reference", "sample" => "$(length(vars)) C5 variables (8 x 4096 -> 4096 samples incl. manikde! of the result), Julia $(VERSION)")))

# C3: mmd of two 1e5-point clouds is 3e10 kernel evaluations; time a bounded row sample
P = 100_000
a = [SVector{3}(randn(3)) for _ in 1:P]; b = [SVector{3}(randn(3)) .+ SA[2.0, 0, 2.0] for _ in 1:P]
Mt = TranslationGroup(3)
rowsN = 2000
mmd(Mt, a[1:100], b[1:100], 100, 100, true; bw = SA[1.0])
t = @elapsed mmd(Mt, a[1:rowsN], b, rowsN, P, true; bw = SA[1.0])
println(JSON3.write(Dict("metric" => "mmd_kernel_pairs_per_s", "value" => (rowsN^2 + rowsN * P + P * P) / t, "unit" => "pairs/s", "cores" => nt,
                         "kind" => "reference", "sample" => "mmd of $(rowsN) x $(P) points (threads = true)")))
