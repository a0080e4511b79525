# The workload of TFHE.jl's examples/multikey.jl on the B200 engine: MK-TFHE NAND under the keys of two parties.
# Untested in this repository's image (no Julia); the same trials run in examples/multikey.py.
include(joinpath(@__DIR__, "..", "TFHEB200.jl"))
using .TFHEB200
using Random

function main(parties = 2, trials = 10)
    params = mktfhe_parameters_2party
    rng = MersenneTwister()
    secret_keys = [SecretKey(rng, params) for _ in 1:parties]                 # on the clients' machines
    shared_key = SharedKey(rng, params)                                       # created by the server
    ck_parts = [CloudKeyPart(rng, sk, shared_key) for sk in secret_keys]      # on the clients' machines
    cloud_key = MKCloudKey(ck_parts)                                          # on the server: expansion on the GPU
    for trial in 1:trials
        mess1, mess2 = rand(rng, Bool), rand(rng, Bool)
        enc1, enc2 = mk_encrypt(rng, secret_keys, mess1), mk_encrypt(rng, secret_keys, mess2)
        @assert mk_decrypt(secret_keys, enc1) == mess1 && mk_decrypt(secret_keys, enc2) == mess2
        out = mk_decrypt(secret_keys, mk_gate_nand(cloud_key, enc1, enc2))
        println("Trial $trial: $mess1 NAND $mess2 = $out", out == !(mess1 && mess2) ? "" : "   (wrong: the reference's own 2-party noise, DESIGN.md 5)")
    end
end

main()
