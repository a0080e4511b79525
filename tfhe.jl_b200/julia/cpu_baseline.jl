# cpu_baseline.jl — times genuine TFHE.jl on the host cores for bench.py --impl reference (used automatically when
# `julia` is on PATH and TFHE_JL_PROJECT points at a TFHE.jl checkout).  NOT RUN in the build image (no Julia).
#
#     julia --threads=auto --project=$TFHE_JL_PROJECT cpu_baseline.jl <gates> <steps> <warmup>
#
# One gate per task, Threads.@threads over the batch (each task owns its ciphertexts; TFHE.jl's transform plans are
# not re-entrant, polynomials.jl:80-103, so threads > 1 needs a TFHE.jl whose plan cache is per-thread — otherwise
# run with --threads=1 and the line reports cores = 1).  Prints one JSON line.
using Random, TFHE
gates, steps, warmup = parse(Int, ARGS[1]), parse(Int, ARGS[2]), parse(Int, ARGS[3])
rng = MersenneTwister(123)
sk, ck = make_key_pair(rng)
x = [encrypt(rng, sk, rand(rng, Bool)) for _ in 1:gates]; y = [encrypt(rng, sk, rand(rng, Bool)) for _ in 1:gates]
out = Vector{Any}(undef, gates)
step() = Threads.@threads for g in 1:gates
    out[g] = gate_nand(ck, x[g], y[g])
end
for _ in 1:warmup; step(); end
t = @elapsed for _ in 1:steps; step(); end
ok = all(decrypt(sk, out[g]) == !(decrypt(sk, x[g]) && decrypt(sk, y[g])) for g in 1:gates)
println("{\"gates_per_s\": $(gates * steps / t), \"ms_per_step\": $(t / steps * 1e3), \"cores\": $(Threads.nthreads()), \"correct\": $(ok)}")
