# TFHEB200.jl — Julia host layer over libtfhe_b200.so (the C ABI of include/tfhe_b200.h).
#
# Drop-in for the API surface of nucypher/TFHE.jl (src/TFHE.jl:24-61): the same exported names, the same
# `LweSample` fields (`params`, `a`, `b`, `current_variance`, lwe.jl:21-29), the same broadcasting behaviour of keys
# and samples (lwe.jl:32, api.jl:103,130), so the documented idiom
#
#     ciphertext1 = encrypt.(Ref(rng), secret_key, bits1)          # docs/src/manual.md:28-35
#     cresult     = gate_xor.(cloud_key, ciphertext1, ciphertext2)
#     result      = decrypt.(secret_key, cresult)
#
# and the `Array{LweSample,1}` signatures of examples/tutorial.jl:42-62 work unchanged — with one difference under the
# hood: a broadcast over vectors of samples is intercepted (`Base.Broadcast.broadcasted`, below), packed into ONE
# Matrix{Int32}(n+1, count) — byte for byte the [count][n+1] batch layout of the C ABI — and evaluated by ONE `ccall`
# (one set of kernel launches for the whole vector, sharded over every GPU of a `CloudKey(...; devices = ...)`),
# instead of one 1.7 ms bootstrap per element.  `LweBatch` exposes that matrix form directly for bulk work.
#
# Julia is not installed in the build image, so this file cannot be executed there; what CAN be checked without Julia is
# checked by tests/test_julia_binding.py: every `ccall` below names a symbol declared in include/tfhe_b200.h with the
# same arity and the same integer widths.  crosscheck.jl compares against genuine TFHE.jl when both are installed.
# There is no CPU fallback: without the library or a GPU every call throws.
#
#     ENV["TFHE_B200_LIB"] = "/path/to/libtfhe_b200.so"   # default: ../libtfhe_b200.so next to this file
module TFHEB200

using Random: AbstractRNG

export make_key_pair, LweSample, LweBatch, SecretKey, CloudKey, SchemeParameters, encrypt, decrypt
export tfhe_parameters_80, tfhe_parameters_128
export gate_nand, gate_or, gate_and, gate_xor, gate_xnor, gate_not, gate_constant, gate_nor
export gate_andny, gate_andyn, gate_orny, gate_oryn, gate_mux
export SharedKey, CloudKeyPart, MKCloudKey, MKLweSample, mk_encrypt, mk_decrypt, mk_gate_nand
export mktfhe_parameters_2party, mktfhe_parameters_4party, mktfhe_parameters_8party

const LIB = get(ENV, "TFHE_B200_LIB", joinpath(@__DIR__, "..", "libtfhe_b200.so"))
const Torus32 = Int32                                           # numeric-functions.jl:1

# ---------------------------------------------------------------------------------------------- C ABI
struct CParams                                                  # tfhe_b200_params
    n::Int32; N::Int32; k::Int32; l::Int32; bgbit::Int32; t::Int32; basebit::Int32; parties::Int32
end

const FLAG_SPLIT_FFT = UInt32(0)
const FLAG_UNSPLIT_FFT = UInt32(1)
@enum GateOp NAND = 0 OR = 1 AND = 2 XOR = 3 XNOR = 4 NOT = 5 CONSTANT = 6 NOR = 7 ANDNY = 8 ANDYN = 9 ORNY = 10 ORYN = 11 MUX = 12

last_error() = unsafe_string(ccall((:tfhe_b200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))
multi_last_error() = unsafe_string(ccall((:tfhe_b200_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))

"One parameter set + one evaluation-key set on one GPU (`multi == false`) or replicated over several (`multi == true`)."
mutable struct Context
    handle::Ptr{Cvoid}
    params::CParams
    multi::Bool
    function Context(p::CParams; device::Integer = 0, devices::Union{Nothing, Vector{<:Integer}, Symbol} = nothing,
                     flags::UInt32 = FLAG_SPLIT_FFT)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        if devices === nothing
            rc = ccall((:tfhe_b200_create, LIB), Cint, (Ref{CParams}, Cint, UInt32, Ref{Ptr{Cvoid}}), p, device, flags, h)
            rc == 0 || error("tfhe_b200_create: " * last_error())
            ctx = new(h[], p, false)
            finalizer(c -> ccall((:tfhe_b200_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
            return ctx
        end
        ids = devices === :all ? Cint[] : Cint.(devices)        # empty list = every visible device
        rc = GC.@preserve ids ccall((:tfhe_b200_multi_create, LIB), Cint, (Ref{CParams}, Ptr{Cint}, Cint, UInt32, Ref{Ptr{Cvoid}}),
                                    p, isempty(ids) ? Ptr{Cint}(C_NULL) : pointer(ids), length(ids), flags, h)
        rc == 0 || error("tfhe_b200_multi_create: " * multi_last_error())
        ctx = new(h[], p, true)
        finalizer(c -> ccall((:tfhe_b200_multi_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
        ctx
    end
end

function check(ctx::Context, rc::Cint)                          # error codes -> Julia exceptions (SURVEY §5)
    rc == 0 && return nothing
    error("tfhe_b200 error $rc: " * (ctx.multi ? multi_last_error() : last_error()))
end

cptr(a::Array{Int32}) = pointer(a)
cptr(::Nothing) = Ptr{Int32}(C_NULL)

# Julia arrays are column-major: Matrix{Int32}(n+1, count) IS the C layout [count][n+1].
function c_gate(ctx::Context, op::GateOp, x, y, z, count::Integer)
    out = Matrix{Int32}(undef, Int(ctx.params.n) * Int(ctx.params.parties) + 1, count)
    GC.@preserve x y z out begin
        rc = if ctx.multi
            ccall((:tfhe_b200_multi_gate_batch, LIB), Cint,
                  (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Csize_t),
                  ctx.handle, Int(op), cptr(x), cptr(y), cptr(z), pointer(out), count)
        else
            ccall((:tfhe_b200_gate_batch, LIB), Cint,
                  (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Csize_t),
                  ctx.handle, Int(op), cptr(x), cptr(y), cptr(z), pointer(out), count)
        end
        check(ctx, rc)
    end
    out
end

"bootstrap (bootstrap.jl:92-95) / mk_bootstrap (mk_internals.jl:512-515) of a batch with an arbitrary test-vector value"
function c_bootstrap(ctx::Context, mu::Torus32, x::Matrix{Int32})
    out = similar(x)
    GC.@preserve x out begin
        rc = if ctx.multi
            ccall((:tfhe_b200_multi_bootstrap_batch, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Csize_t),
                  ctx.handle, mu, pointer(x), pointer(out), size(x, 2))
        elseif ctx.params.parties == 1
            ccall((:tfhe_b200_bootstrap_batch, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Csize_t),
                  ctx.handle, mu, pointer(x), pointer(out), size(x, 2))
        else
            ccall((:tfhe_b200_mk_bootstrap_batch, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Csize_t),
                  ctx.handle, mu, pointer(x), pointer(out), size(x, 2))
        end
        check(ctx, rc)
    end
    out
end

function c_polymul(ctx::Context, x::Matrix{Int32}, y::Matrix{Int32})     # transformed_mul, polynomials.jl:142-144
    @assert !ctx.multi
    out = similar(x)
    GC.@preserve x y out check(ctx, ccall((:tfhe_b200_polymul_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Csize_t), ctx.handle, pointer(x), pointer(y), pointer(out), size(x, 2)))
    out
end

function load_bk!(ctx::Context, bk::Array{Int32})
    GC.@preserve bk begin
        rc = if ctx.multi && ctx.params.parties == 1
            ccall((:tfhe_b200_multi_load_bk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, pointer(bk))
        elseif ctx.multi
            ccall((:tfhe_b200_multi_mk_load_bk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, pointer(bk))
        elseif ctx.params.parties == 1
            ccall((:tfhe_b200_load_bk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, pointer(bk))
        else
            ccall((:tfhe_b200_mk_load_bk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, pointer(bk))
        end
        check(ctx, rc)
    end
end

function load_ksk!(ctx::Context, ksk::Array{Int32})
    GC.@preserve ksk begin
        rc = if ctx.multi && ctx.params.parties == 1
            ccall((:tfhe_b200_multi_load_ksk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, pointer(ksk))
        elseif ctx.multi
            ccall((:tfhe_b200_multi_mk_load_ksk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, pointer(ksk))
        elseif ctx.params.parties == 1
            ccall((:tfhe_b200_load_ksk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, pointer(ksk))
        else
            ccall((:tfhe_b200_mk_load_ksk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, pointer(ksk))
        end
        check(ctx, rc)
    end
end

# ---------------------------------------------------------------------------------------------- numeric-functions.jl
rand_uniform_bool(rng::AbstractRNG, dims...) = rand(rng, Int32(0):Int32(1), dims...)
rand_uniform_torus32(rng::AbstractRNG, dims...) = rand(rng, Torus32, dims...)
dtot32(d::Float64) = trunc(Int32, d * 2^32)                    # numeric-functions.jl:51-53 (InexactError outside [-0.5, 0.5))
rand_gaussian_torus32(rng::AbstractRNG, sigma::Float64, dims...) = dtot32.(randn(rng, dims...) .* sigma)
encode_message(mu::Int, message_space::Int) = Torus32(mu) << (32 - trailing_zeros(message_space))

# ---------------------------------------------------------------------------------------------- api.jl
struct SchemeParameters                                         # api.jl:4-21
    lwe_size::Int; lwe_noise_stddev::Float64
    tlwe_polynomial_degree::Int; tlwe_mask_size::Int
    bs_decomp_length::Int; bs_log2_base::Int; bs_noise_stddev::Float64
    ks_decomp_length::Int; ks_log2_base::Int; ks_noise_stddev::Float64
    max_parties::Int
end

tfhe_parameters_80(; tlwe_mask_size::Int = 1) = SchemeParameters(500, 1 / 2^15 * sqrt(2 / pi), 1024, tlwe_mask_size,
    2, 10, 9e-9 * sqrt(2 / pi), 8, 2, 1 / 2^15 * sqrt(2 / pi), 1)                    # api.jl:30-45
tfhe_parameters_128(; tlwe_mask_size::Int = 1) = SchemeParameters(630, 1 / 2^15, 1024, tlwe_mask_size,
    3, 7, 1 / 2^25, 8, 2, 1 / 2^15, 1)                                               # api.jl:55-69
const mktfhe_parameters_2party = SchemeParameters(500, 0.012467, 1024, 1, 4, 7, 3.29e-10, 8, 2, 2.44e-5, 2)   # mk_api.jl:4-10
const mktfhe_parameters_4party = SchemeParameters(500, 0.012467, 1024, 1, 5, 6, 3.29e-10, 8, 2, 2.44e-5, 4)   # mk_api.jl:16-22
const mktfhe_parameters_8party = SchemeParameters(500, 0.012467, 1024, 1, 8, 4, 3.29e-10, 8, 2, 2.44e-5, 8)   # mk_api.jl:28-34

cparams(p::SchemeParameters, parties::Int) = CParams(p.lwe_size, p.tlwe_polynomial_degree, p.tlwe_mask_size,
    p.bs_decomp_length, p.bs_log2_base, p.ks_decomp_length, p.ks_log2_base, parties)

struct LweParams                                                # lwe.jl:1-8
    size::Int
end

"An encrypted bit: the reference's structure, field for field (lwe.jl:21-29)."
mutable struct LweSample
    params::LweParams
    a::Array{Torus32, 1}
    b::Torus32
    current_variance::Float64
end
Base.Broadcast.broadcastable(lwe::LweSample) = (lwe,)           # lwe.jl:32
Base.:+(x::LweSample, y::LweSample) = LweSample(x.params, x.a .+ y.a, x.b + y.b, x.current_variance + y.current_variance)   # lwe.jl:67-68
Base.:-(x::LweSample, y::LweSample) = LweSample(x.params, x.a .- y.a, x.b - y.b, x.current_variance + y.current_variance)   # lwe.jl:71-72
Base.:-(x::LweSample) = LweSample(x.params, .-x.a, -x.b, x.current_variance)                                                # lwe.jl:74
Base.:*(x::LweSample, y::Integer) = LweSample(x.params, x.a .* Int32(y), x.b * Int32(y), x.current_variance * y^2)           # lwe.jl:77-82
Base.:*(y::Integer, x::LweSample) = x * y

"""
A batch of encrypted bits in the C ABI's layout: `data` is (n+1, count), one ciphertext per column (rows 1:n = a,
row n+1 = b).  `LweBatch(v)` packs a vector of samples, `collect(batch)` / `batch[i]` unpack.  All `gate_*` functions
accept batches and evaluate them with one call into the library.
"""
struct LweBatch
    data::Matrix{Int32}
    current_variance::Float64
end
LweBatch(v::AbstractVector{LweSample}) =
    LweBatch(isempty(v) ? Matrix{Int32}(undef, 1, 0) : hcat((vcat(s.a, s.b) for s in v)...), isempty(v) ? 0.0 : maximum(s.current_variance for s in v))
Base.length(x::LweBatch) = size(x.data, 2)
Base.getindex(x::LweBatch, i::Integer) = LweSample(LweParams(size(x.data, 1) - 1), x.data[1:end-1, i], x.data[end, i], x.current_variance)
Base.collect(x::LweBatch) = [x[i] for i in 1:length(x)]

struct SecretKey                                                # api.jl:92-100
    params::SchemeParameters
    key::Vector{Int32}
    SecretKey(rng::AbstractRNG, params::SchemeParameters) = new(params, rand_uniform_bool(rng, params.lwe_size))
end
Base.Broadcast.broadcastable(sk::SecretKey) = (sk,)             # api.jl:103

# tlwe.jl:63-73 for `count` samples at once; the products S (*) a run on the GPU (kernel K1).
# All sums stay in Int32 (wrap-around mod 2^32, as in the reference): `sum` over Int32 would widen to Int64.
function tlwe_encrypt_zero(rng, ctx::Context, alpha::Float64, tlwe_key::Matrix{Int32}, count::Int)
    N, k = size(tlwe_key)
    a = rand_uniform_torus32(rng, N, k, count)
    b = rand_gaussian_torus32(rng, alpha, N, count)
    keys = repeat(reshape(tlwe_key, N, k, 1), 1, 1, count)
    prod = reshape(c_polymul(ctx, reshape(keys, N, k * count), reshape(a, N, k * count)), N, k, count)
    for c in 1:k
        b .+= view(prod, :, c, :)
    end
    cat(a, reshape(b, N, 1, count), dims = 2)                   # (N, k+1, count)
end

# bootstrap.jl:6-15 / tgsw.jl:52-88 in coefficient form; memory order [n][l][k+1][k+1][N] of the C ABI
function bootstrap_key(rng, ctx, alpha, lwe_key::Vector{Int32}, tlwe_key::Matrix{Int32}, l::Int, bgbit::Int)
    N, k = size(tlwe_key); n = length(lwe_key)
    bk = reshape(tlwe_encrypt_zero(rng, ctx, alpha, tlwe_key, n * l * (k + 1)), N, k + 1, k + 1, l, n)   # (N, c, j, r, i)
    for i in 1:n, r in 1:l, j in 1:(k + 1)
        bk[1, j, j, r, i] += lwe_key[i] * (Int32(1) << (32 - r * bgbit))                                 # tgsw.jl:62-69
    end
    bk
end

wrapdot(a::AbstractVector{Int32}, key::AbstractVector{Int32}) = reduce(+, a .* key; init = Int32(0))      # Int32 wrap-around

# keyswitch.jl:14-41; memory order [N*k][t][base-1][n+1]
function keyswitch_key(rng, alpha, t::Int, basebit::Int, out_key::Vector{Int32}, in_key::Vector{Int32})
    base = 1 << basebit; n = length(out_key); Nk = length(in_key)
    noise = randn(rng, base - 1, t, Nk) .* alpha
    noise .-= sum(noise) / length(noise)                                                                 # keyswitch.jl:29
    ks = Array{Int32}(undef, n + 1, base - 1, t, Nk)
    for i in 1:Nk, j in 1:t, h in 1:(base - 1)
        a = rand_uniform_torus32(rng, n)
        message = (in_key[i] * Int32(h)) << (32 - j * basebit)                                           # keyswitch.jl:35
        ks[1:n, h, j, i] = a
        ks[n + 1, h, j, i] = message + dtot32(noise[h, j, i]) + wrapdot(a, out_key)                      # lwe.jl:49-55
    end
    ks
end

"""
    CloudKey(rng, secret_key; device = 0, devices = nothing, flags = FLAG_SPLIT_FFT)

api.jl:111-127.  `devices = :all` or a vector of GPU ordinals replicates the key on those GPUs; every gate call then
shards its batch across them (tfhe_b200_multi_*).  The int32 coefficient form of the bootstrapping key is kept
(the reference keeps only its transform, bootstrap.jl:12-14).
"""
struct CloudKey
    params::SchemeParameters
    ctx::Context
    bootstrap_key::Array{Int32}
    keyswitch_key::Array{Int32}
    function CloudKey(rng::AbstractRNG, secret_key::SecretKey; device::Integer = 0, devices = nothing, flags::UInt32 = FLAG_SPLIT_FFT)
        p = secret_key.params
        kctx = Context(cparams(p, 1); device = devices === nothing || devices === :all ? device : first(devices), flags = flags)
        tlwe_key = rand_uniform_bool(rng, p.tlwe_polynomial_degree, p.tlwe_mask_size)
        bk = bootstrap_key(rng, kctx, p.bs_noise_stddev, secret_key.key, tlwe_key, p.bs_decomp_length, p.bs_log2_base)
        ks = keyswitch_key(rng, p.ks_noise_stddev, p.ks_decomp_length, p.ks_log2_base, secret_key.key, vec(tlwe_key))
        ctx = devices === nothing ? kctx : Context(cparams(p, 1); devices = devices, flags = flags)
        load_bk!(ctx, bk); load_ksk!(ctx, ks)
        new(p, ctx, bk, ks)
    end
end
Base.Broadcast.broadcastable(ck::CloudKey) = (ck,)              # api.jl:130

function make_key_pair(rng::AbstractRNG, params::Union{Nothing, SchemeParameters} = nothing; kwargs...)   # api.jl:139-146
    params === nothing && (params = tfhe_parameters_80())
    secret_key = SecretKey(rng, params)
    secret_key, CloudKey(rng, secret_key; kwargs...)
end

function lwe_encrypt(rng, message::Torus32, alpha::Float64, key::Vector{Int32})                           # lwe.jl:38-43
    a = rand_uniform_torus32(rng, length(key))
    LweSample(LweParams(length(key)), a, message + dtot32(randn(rng) * alpha) + wrapdot(a, key), alpha^2)
end

"api.jl:155-158"
encrypt(rng::AbstractRNG, key::SecretKey, message::Bool) =
    lwe_encrypt(rng, encode_message(message ? 1 : -1, 8), key.params.lwe_noise_stddev, key.key)
"Batched form: one ciphertext per column of an `LweBatch`."
encrypt(rng::AbstractRNG, key::SecretKey, messages::AbstractVector{Bool}) = LweBatch([encrypt(rng, key, m) for m in messages])

lwe_phase(a::AbstractVector{Int32}, b::Int32, key) = b - wrapdot(a, key)                                   # lwe.jl:59
decrypt(key::SecretKey, sample::LweSample) = lwe_phase(sample.a, sample.b, key.key) > 0                   # api.jl:167-169
decrypt(key::SecretKey, batch::LweBatch) =
    [lwe_phase(view(batch.data, 1:size(batch.data, 1) - 1, g), batch.data[end, g], key.key) > 0 for g in 1:length(batch)]

# ---------------------------------------------------------------------------------------------- gates.jl
as_matrix(x::LweSample) = reshape(vcat(x.a, x.b), :, 1)
as_matrix(x::LweBatch) = x.data
as_matrix(x::AbstractVector{LweSample}) = LweBatch(x).data
unpack(out::Matrix{Int32}, ::LweSample) = LweBatch(out, 0.0)[1]
unpack(out::Matrix{Int32}, ::LweBatch) = LweBatch(out, 0.0)
unpack(out::Matrix{Int32}, ::AbstractVector{LweSample}) = collect(LweBatch(out, 0.0))

"One library call for the whole batch, whatever form the operands have (a sample, an `LweBatch`, a vector of samples)."
function gate(ck::CloudKey, op::GateOp, xs...)
    mats = map(as_matrix, xs)
    count = size(mats[1], 2)
    all(m -> size(m, 2) == count, mats) || throw(DimensionMismatch("gate operands hold different numbers of ciphertexts"))
    out = c_gate(ck.ctx, op, mats[1], length(mats) > 1 ? mats[2] : nothing, length(mats) > 2 ? mats[3] : nothing, count)
    unpack(out, xs[1])
end
gate_nand(ck::CloudKey, x, y) = gate(ck, NAND, x, y)          # gates.jl:15-18
gate_or(ck::CloudKey, x, y) = gate(ck, OR, x, y)              # gates.jl:27-30
gate_and(ck::CloudKey, x, y) = gate(ck, AND, x, y)            # gates.jl:39-42
gate_xor(ck::CloudKey, x, y) = gate(ck, XOR, x, y)            # gates.jl:51-54
gate_xnor(ck::CloudKey, x, y) = gate(ck, XNOR, x, y)          # gates.jl:63-66
gate_not(ck::CloudKey, x) = gate(ck, NOT, x)                  # gates.jl:76-79
gate_nor(ck::CloudKey, x, y) = gate(ck, NOR, x, y)            # gates.jl:102-105
gate_andny(ck::CloudKey, x, y) = gate(ck, ANDNY, x, y)        # gates.jl:114-117
gate_andyn(ck::CloudKey, x, y) = gate(ck, ANDYN, x, y)        # gates.jl:126-129
gate_orny(ck::CloudKey, x, y) = gate(ck, ORNY, x, y)          # gates.jl:138-141
gate_oryn(ck::CloudKey, x, y) = gate(ck, ORYN, x, y)          # gates.jl:150-153
gate_mux(ck::CloudKey, x, y, z) = gate(ck, MUX, x, y, z)      # gates.jl:163-177
function gate_constant(ck::CloudKey, value::Bool)             # gates.jl:91-93
    flags = zeros(Int32, ck.params.lwe_size + 1, 1); flags[1, 1] = value
    LweBatch(c_gate(ck.ctx, CONSTANT, flags, nothing, nothing, 1), 0.0)[1]
end

# Broadcasting.  `gate_xor.(cloud_key, xs, ys)` over vectors of samples lowers to
# `materialize(broadcasted(gate_xor, cloud_key, xs, ys))`; these methods return the finished Vector{LweSample}
# (materialize of an Array is the identity), computed as ONE batch instead of element by element.  Scalars mixed with
# vectors (a single sample against a vector) are repeated to the common length, as broadcasting would do.
const SampleVec = AbstractVector{LweSample}
repeat_to(x::LweSample, count::Int) = repeat(as_matrix(x), 1, count)
repeat_to(x::SampleVec, count::Int) = (length(x) == count || throw(DimensionMismatch("broadcast over ciphertext vectors of different lengths")); as_matrix(x))
function broadcast_gate(ck::CloudKey, op::GateOp, xs...)
    count = maximum(x isa LweSample ? 1 : length(x) for x in xs)
    mats = map(x -> repeat_to(x, count), xs)
    out = c_gate(ck.ctx, op, mats[1], length(mats) > 1 ? mats[2] : nothing, length(mats) > 2 ? mats[3] : nothing, count)
    collect(LweBatch(out, 0.0))
end
for (f, op) in ((:gate_nand, NAND), (:gate_or, OR), (:gate_and, AND), (:gate_xor, XOR), (:gate_xnor, XNOR), (:gate_nor, NOR),
                (:gate_andny, ANDNY), (:gate_andyn, ANDYN), (:gate_orny, ORNY), (:gate_oryn, ORYN))
    @eval begin
        Base.Broadcast.broadcasted(::typeof($f), ck::CloudKey, x::SampleVec, y::SampleVec) = broadcast_gate(ck, $op, x, y)
        Base.Broadcast.broadcasted(::typeof($f), ck::CloudKey, x::LweSample, y::SampleVec) = broadcast_gate(ck, $op, x, y)
        Base.Broadcast.broadcasted(::typeof($f), ck::CloudKey, x::SampleVec, y::LweSample) = broadcast_gate(ck, $op, x, y)
    end
end
Base.Broadcast.broadcasted(::typeof(gate_not), ck::CloudKey, x::SampleVec) = broadcast_gate(ck, NOT, x)
Base.Broadcast.broadcasted(::typeof(gate_mux), ck::CloudKey, x::Union{LweSample, SampleVec}, y::Union{LweSample, SampleVec},
                           z::SampleVec) = broadcast_gate(ck, MUX, x, y, z)
Base.Broadcast.broadcasted(::typeof(gate_mux), ck::CloudKey, x::SampleVec, y::Union{LweSample, SampleVec}, z::LweSample) =
    broadcast_gate(ck, MUX, x, y, z)
Base.Broadcast.broadcasted(::typeof(gate_mux), ck::CloudKey, x::LweSample, y::SampleVec, z::LweSample) = broadcast_gate(ck, MUX, x, y, z)
# encrypt.(Ref(rng), secret_key, bits) and decrypt.(secret_key, samples): vectorised on the host in one pass
Base.Broadcast.broadcasted(::typeof(encrypt), rng::Base.RefValue{<:AbstractRNG}, key::SecretKey, bits::AbstractVector{Bool}) =
    [encrypt(rng[], key, m) for m in bits]
Base.Broadcast.broadcasted(::typeof(decrypt), key::SecretKey, xs::SampleVec) = decrypt(key, LweBatch(xs))

# ---------------------------------------------------------------------------------------------- multi-key
"mk_internals.jl:6-18: `data` is (p*n+1) or (p*n+1, count): a[:, party] blocks then the joint b."
struct MKLweSample
    data::Array{Int32}
    parties::Int
    current_variance::Float64
end
Base.Broadcast.broadcastable(x::MKLweSample) = (x,)

struct SharedKey                                                # mk_internals.jl:101-112, mk_api.jl:44-50
    params::SchemeParameters
    a::Matrix{Int32}                                            # (N, l)
    SharedKey(rng::AbstractRNG, params::SchemeParameters) =
        new(params, rand_uniform_torus32(rng, params.tlwe_polynomial_degree, params.bs_decomp_length))
end

const keygen_ctx = Dict{Tuple{SchemeParameters, Int}, Context}()
mk_keygen_ctx(p::SchemeParameters, device::Int) = get!(() -> Context(cparams(p, p.max_parties); device = device), keygen_ctx, (p, device))

mulpoly(ctx, x::Array{Int32}, y::Array{Int32}) = reshape(c_polymul(ctx, reshape(x, size(x, 1), :), reshape(y, size(y, 1), :)), size(y))

struct CloudKeyPart                                             # mk_api.jl:61-77
    params::SchemeParameters
    public_b::Matrix{Int32}                                     # PublicKey.b, (N, l)         mk_internals.jl:115-139
    uni_enc::Dict{Symbol, Array{Int32, 3}}                      # c0,c1,d0,d1,f0,f1: (N, l, n)  mk_internals.jl:185-227
    ks::Array{Int32}
    device::Int
    function CloudKeyPart(rng, secret_key::SecretKey, shared_key::SharedKey; device::Integer = 0)
        p = secret_key.params
        ctx = mk_keygen_ctx(p, Int(device))
        N, l, n, alpha = p.tlwe_polynomial_degree, p.bs_decomp_length, p.lwe_size, p.bs_noise_stddev
        S = rand_uniform_bool(rng, N)
        rep(v, dims...) = repeat(reshape(v, N, ntuple(_ -> 1, length(dims))...), 1, dims...)
        public_b = mulpoly(ctx, rep(S, l), shared_key.a) .+ rand_gaussian_torus32(rng, alpha, N, l)
        gadget = [Int32(1) << (32 - r * p.bs_log2_base) for r in 1:l]
        r = rand_uniform_bool(rng, N, 1, n)
        rr = repeat(r, 1, l, 1)
        c1 = rand_uniform_torus32(rng, N, l, n)
        c0 = mulpoly(ctx, rep(S, l, n), c1) .+ rand_gaussian_torus32(rng, alpha, N, l, n)
        d1 = mulpoly(ctx, rr, repeat(shared_key.a, 1, 1, n)) .+ rand_gaussian_torus32(rng, alpha, N, l, n)
        d0 = mulpoly(ctx, rr, repeat(public_b, 1, 1, n)) .+ rand_gaussian_torus32(rng, alpha, N, l, n)
        f1 = rand_uniform_torus32(rng, N, l, n)
        f0 = mulpoly(ctx, rep(S, l, n), f1) .+ rand_gaussian_torus32(rng, alpha, N, l, n)
        for j in 1:n, i in 1:l
            c0[1, i, j] += secret_key.key[j] * gadget[i]                                                 # :200-204
            d1[1, i, j] += secret_key.key[j] * gadget[i]                                                 # :207-211
            f0[:, i, j] .+= r[:, 1, j] .* gadget[i]                                                      # :220-224
        end
        ks = keyswitch_key(rng, p.ks_noise_stddev, p.ks_decomp_length, p.ks_log2_base, secret_key.key, S)
        new(p, public_b, Dict(:c0 => c0, :c1 => c1, :d0 => d0, :d1 => d1, :f0 => f0, :f1 => f1), ks, Int(device))
    end
end

function decompose(x::Array{Int32}, l::Int, bgbit::Int)         # tgsw.jl:99-117 (host side, key expansion only)
    offset = reinterpret(Int32, UInt32((sum(Int64(1) << (32 - r * bgbit) for r in 1:l) * (1 << (bgbit - 1))) % 2^32))
    [((x .+ offset) .>> (32 - r * bgbit)) .& Int32((1 << bgbit) - 1) .- Int32(1 << (bgbit - 1)) for r in 1:l]
end

struct MKCloudKey                                               # mk_api.jl:85-101
    parties::Int
    params::SchemeParameters
    ctx::Context
    function MKCloudKey(ck_parts::Vector{CloudKeyPart}; devices = nothing, flags::UInt32 = FLAG_SPLIT_FFT)
        params = ck_parts[1].params; p = length(ck_parts)
        @assert p <= params.max_parties                                                                  # mk_api.jl:94
        kctx = mk_keygen_ctx(params, ck_parts[1].device)
        N, l, n = params.tlwe_polynomial_degree, params.bs_decomp_length, params.lwe_size
        # C-ABI order [party][n][x(l,p) | y(l,p) | c0(l) | c1(l)][N]  ==  Julia (N, l*(2p+2), n, party)
        bk = Array{Int32}(undef, N, l * (2p + 2), n, p)
        for (i, part) in enumerate(ck_parts)                                                             # RGSW.Expand, mk_internals.jl:304-345
            ue = part.uni_enc
            for jj in 1:l, ii in 1:p
                xi = (jj - 1) * p + ii; yi = l * p + xi
                bk[:, xi, :, i] = ue[:d0][:, jj, :]                                                      # :327
                if ii == i
                    bk[:, yi, :, i] = ue[:d1][:, jj, :]                                                  # :336
                else
                    u = decompose(ck_parts[ii].public_b[:, jj] .- part.public_b[:, jj], l, params.bs_log2_base)   # :321
                    accx = zeros(Int32, N, n); accy = zeros(Int32, N, n)
                    for r in 1:l
                        ur = repeat(u[r], 1, n)
                        accx .+= c_polymul(kctx, ur, ue[:f0][:, r, :])                                   # :330
                        accy .+= c_polymul(kctx, ur, ue[:f1][:, r, :])                                   # :338
                    end
                    bk[:, xi, :, i] .+= accx
                    bk[:, yi, :, i] = accy
                end
            end
            bk[:, (2l * p + 1):(2l * p + l), :, i] = ue[:c0]
            bk[:, (2l * p + l + 1):(2l * p + 2l), :, i] = ue[:c1]
        end
        ctx = devices === nothing ? Context(cparams(params, p); device = ck_parts[1].device, flags = flags) :
                                    Context(cparams(params, p); devices = devices, flags = flags)
        load_bk!(ctx, bk)
        load_ksk!(ctx, cat([part.ks for part in ck_parts]..., dims = 5))
        new(p, params, ctx)
    end
end
Base.Broadcast.broadcastable(ck::MKCloudKey) = (ck,)

function mk_encrypt(rng, secret_keys::Vector{SecretKey}, message::Bool)                                   # mk_api.jl:110-126
    params = secret_keys[1].params
    keys = vcat([sk.key for sk in secret_keys]...)
    a = rand_uniform_torus32(rng, length(keys))
    b = encode_message(message ? 1 : -1, 8) + dtot32(randn(rng) * params.lwe_noise_stddev) + wrapdot(a, keys)
    MKLweSample(vcat(a, b), length(secret_keys), params.lwe_noise_stddev^2)
end
mk_encrypt(rng, secret_keys::Vector{SecretKey}, messages::AbstractVector{Bool}) =
    MKLweSample(hcat([mk_encrypt(rng, secret_keys, m).data for m in messages]...), length(secret_keys), secret_keys[1].params.lwe_noise_stddev^2)

mk_phase(data::AbstractVector{Int32}, keys) = data[end] - wrapdot(view(data, 1:length(data) - 1), keys)
function mk_decrypt(secret_keys::Vector{SecretKey}, sample::MKLweSample)                                  # mk_api.jl:135-138
    keys = vcat([sk.key for sk in secret_keys]...)
    ndims(sample.data) == 1 ? mk_phase(sample.data, keys) > 0 : [mk_phase(view(sample.data, :, g), keys) > 0 for g in 1:size(sample.data, 2)]
end

mk_matrix(x::MKLweSample) = ndims(x.data) == 1 ? reshape(x.data, :, 1) : x.data

function mk_gate_nand(ck::MKCloudKey, x::MKLweSample, y::MKLweSample)                                      # mk_gates.jl:7-12
    xm = mk_matrix(x); ym = mk_matrix(y)
    out = similar(xm)
    GC.@preserve xm ym out begin
        rc = if ck.ctx.multi
            ccall((:tfhe_b200_multi_mk_nand_batch, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Csize_t),
                  ck.ctx.handle, pointer(xm), pointer(ym), pointer(out), size(xm, 2))
        else
            ccall((:tfhe_b200_mk_nand_batch, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Csize_t),
                  ck.ctx.handle, pointer(xm), pointer(ym), pointer(out), size(xm, 2))
        end
        check(ck.ctx, rc)
    end
    MKLweSample(ndims(x.data) == 1 ? vec(out) : out, ck.parties, 0.0)
end

"mk_bootstrap (mk_internals.jl:512-515) with an arbitrary test-vector value `mu`"
mk_bootstrap(ck::MKCloudKey, mu::Torus32, x::MKLweSample) =
    (out = c_bootstrap(ck.ctx, mu, mk_matrix(x)); MKLweSample(ndims(x.data) == 1 ? vec(out) : out, ck.parties, 0.0))

"bootstrap (bootstrap.jl:92-95) with an arbitrary test-vector value `mu`"
bootstrap(ck::CloudKey, mu::Torus32, x::Union{LweSample, LweBatch}) = unpack(c_bootstrap(ck.ctx, mu, as_matrix(x)), x)

# mk_gate_nand.(ck, xs, ys) over vectors of single MK samples: one call
function Base.Broadcast.broadcasted(::typeof(mk_gate_nand), ck::MKCloudKey, xs::AbstractVector{MKLweSample}, ys::AbstractVector{MKLweSample})
    length(xs) == length(ys) || throw(DimensionMismatch("broadcast over ciphertext vectors of different lengths"))
    x = MKLweSample(hcat((vec(s.data) for s in xs)...), ck.parties, 0.0)
    y = MKLweSample(hcat((vec(s.data) for s in ys)...), ck.parties, 0.0)
    out = mk_gate_nand(ck, x, y)
    [MKLweSample(out.data[:, g], ck.parties, 0.0) for g in 1:size(out.data, 2)]
end

end # module
