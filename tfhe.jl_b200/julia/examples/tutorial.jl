# The workload of TFHE.jl's examples/tutorial.jl on the B200 engine: the minimum of two encrypted 16-bit integers.
# Only the package changes (`using .TFHEB200` instead of `using TFHE`); the gate calls are the reference's own.
# Untested in this repository's image (no Julia); the same circuit runs in examples/tutorial.py.
#
#   julia tfhe.jl_b200/julia/examples/tutorial.jl
include(joinpath(@__DIR__, "..", "TFHEB200.jl"))
using .TFHEB200
using Random

bits_of(x::UInt16) = [((x >> (i - 1)) & 1) != 0 for i in 1:16]
value_of(bits) = reduce(|, (UInt16(b) << (i - 1) for (i, b) in enumerate(bits)); init = UInt16(0))

# comparator chain from the least significant bit: carry = (a < b so far); then the smaller operand bit by bit
function encrypted_minimum(ck::CloudKey, a::Vector{LweSample}, b::Vector{LweSample})
    carry = gate_constant(ck, false)
    for i in eachindex(a)
        carry = gate_mux(ck, gate_xnor(ck, a[i], b[i]), carry, a[i])
    end
    gate_mux.(ck, carry, b, a)            # 16 independent MUXes: ONE library call (broadcast interception)
end

rng = MersenneTwister(123)
secret_key, cloud_key = make_key_pair(rng)
ciphertext1 = encrypt.(Ref(rng), secret_key, bits_of(UInt16(2017)))
ciphertext2 = encrypt.(Ref(rng), secret_key, bits_of(UInt16(42)))
answer = encrypted_minimum(cloud_key, ciphertext1, ciphertext2)
println("Answer: ", value_of(decrypt.(secret_key, answer)))
