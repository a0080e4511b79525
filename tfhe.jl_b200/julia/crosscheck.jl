# crosscheck.jl — pins libtfhe_b200.so against the REAL TFHE.jl (SURVEY.md §8c: ciphertext-level parity with
# TFHE.jl itself cannot be pinned in the build image because Julia is not installed there).
#
#     julia --project=/path/to/TFHE.jl tfhe.jl_b200/julia/crosscheck.jl [count [fixture_dir]]
#
# Generates keys and ciphertexts with TFHE.jl (MersenneTwister(123), as test/runtests.jl:27), exports the keys in
# the C ABI's int32 layouts, runs every bootstrapped gate both through TFHE.jl and through the library on the
# SAME ciphertexts and requires bit-identical LWE samples.  NOT RUN in the build image (no Julia); every library
# call below is one the Python tests bind.
using Random
using TFHE
include(joinpath(@__DIR__, "TFHEB200.jl"))
const B = TFHEB200

count = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 8
rng = MersenneTwister(123)
sk, ck = TFHE.make_key_pair(rng)
p = ck.params
n, N, k, l = p.lwe_size, p.tlwe_polynomial_degree, p.tlwe_mask_size, p.bs_decomp_length

# bootstrap key: TFHE.jl keeps only the transformed form (bootstrap.jl:12-14); the int32 coefficients are
# recovered exactly by inverse_transform (SURVEY.md App. B).  C layout [n][l][row][component][N] = Julia (N, c, j, r, i).
bk = Array{Int32}(undef, N, k + 1, k + 1, l, n)
for i in 1:n, r in 1:l, j in 1:(k + 1), c in 1:(k + 1)
    bk[:, c, j, r, i] = TFHE.inverse_transform(ck.bootstrap_key.key[i].samples[r, j].a[c]).coeffs
end
# keyswitch key: key[h, j, i] (keyswitch.jl:36-38) -> (n+1, h, j, i)
t, base1 = p.ks_decomp_length, (1 << p.ks_log2_base) - 1
ks = Array{Int32}(undef, n + 1, base1, t, N * k)
for i in 1:(N * k), j in 1:t, h in 1:base1
    s = ck.keyswitch_key.key[h, j, i]
    ks[1:n, h, j, i] = s.a; ks[n + 1, h, j, i] = s.b
end

ctx = B.Context(B.cparams(B.tfhe_parameters_80(), 1))
B.load_bk!(ctx, bk); B.load_ksk!(ctx, ks)

tomat(cts) = hcat([vcat(ct.a, ct.b) for ct in cts]...)
bits = rand(rng, Bool, count, 3)
x, y, z = ([TFHE.encrypt(rng, sk, bits[g, c]) for g in 1:count] for c in 1:3)
binary = [(B.NAND, TFHE.gate_nand), (B.OR, TFHE.gate_or), (B.AND, TFHE.gate_and), (B.XOR, TFHE.gate_xor),
          (B.XNOR, TFHE.gate_xnor), (B.NOR, TFHE.gate_nor), (B.ANDNY, TFHE.gate_andny), (B.ANDYN, TFHE.gate_andyn),
          (B.ORNY, TFHE.gate_orny), (B.ORYN, TFHE.gate_oryn)]
ok = true
for (op, ref) in binary
    want = tomat([ref(ck, x[g], y[g]) for g in 1:count])
    got = B.c_gate(ctx, op, tomat(x), tomat(y), nothing, count)
    same = want == got
    println(rpad(string(op), 6), same ? " identical" : " MISMATCH")
    global ok &= same
end
want = tomat([TFHE.gate_mux(ck, x[g], y[g], z[g]) for g in 1:count])
got = B.c_gate(ctx, B.MUX, tomat(x), tomat(y), tomat(z), count)
println("MUX    ", want == got ? "identical" : "MISMATCH"); ok &= want == got

# Optional second argument: a directory to receive reference-pinned fixtures (raw little-endian Int32 files + a manifest)
# for tests/test_reference_fixtures.py — commit them as tests/golden/tfhejl/ and the oracle is pinned to TFHE.jl itself.
if length(ARGS) > 1
    dir = ARGS[2]; mkpath(dir)
    dump(name, a) = open(io -> write(io, Array{Int32}(a)), joinpath(dir, name * ".bin"), "w")
    dump("lwe_key", sk.key.key); dump("bk", bk); dump("ksk", ks)
    dump("x", tomat(x)); dump("y", tomat(y)); dump("z", tomat(z))
    for (op, ref) in binary
        dump("out_" * string(op), tomat([ref(ck, x[g], y[g]) for g in 1:count]))
    end
    dump("out_MUX", want)
    open(joinpath(dir, "manifest.json"), "w") do io
        print(io, "{\"source\": \"TFHE.jl via crosscheck.jl, MersenneTwister(123)\", \"count\": $count, \"n\": $n, \"N\": $N, \"k\": $k, \"l\": $l, ",
              "\"bgbit\": $(p.bs_log2_base), \"t\": $t, \"basebit\": $(p.ks_log2_base), \"gates\": [",
              join(["\"" * string(op) * "\"" for (op, _) in binary], ", "), ", \"MUX\"]}")
    end
end
exit(ok ? 0 : 1)
