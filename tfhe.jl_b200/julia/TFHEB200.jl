# TFHEB200.jl — Julia host layer over libtfhe_b200.so (the C ABI of include/tfhe_b200.h).
#
# Keeps the API surface of nucypher/TFHE.jl (src/TFHE.jl:24-61): make_key_pair, encrypt, decrypt, the
# thirteen gate_* functions and the MK entry points, and adds batched array variants: an `LweSample` may
# hold a Matrix{Int32}(n+1, count) (one ciphertext per column — exactly the [count][n+1] layout of the C
# ABI), so `gate_nand(ck, x, y)` on batches is ONE `ccall` and one set of kernel launches.
#
# This file is a 1:1 thin mirror of tfhe.jl_b200/api.py + _cabi.py (which the test-suite exercises through
# ctypes).  Julia is not installed in the build image, so this wrapper is shipped untested; every call it
# makes is a symbol the tests bind.  There is no CPU fallback: without the library / a GPU, calls throw.
#
#     ENV["TFHE_B200_LIB"] = "/path/to/libtfhe_b200.so"   # default: ../libtfhe_b200.so next to this file
#     using .TFHEB200, Random
#     sk, ck = make_key_pair(MersenneTwister(123))
#     x = encrypt(rng, sk, rand(Bool, 65536)); y = encrypt(rng, sk, rand(Bool, 65536))
#     z = gate_nand(ck, x, y)            # 65 536 bootstrapped NANDs on the B200
#     decrypt(sk, z)
module TFHEB200

using Random: AbstractRNG

export make_key_pair, LweSample, SecretKey, CloudKey, encrypt, decrypt, tfhe_parameters_80, tfhe_parameters_128
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

mutable struct Context
    handle::Ptr{Cvoid}
    params::CParams
    function Context(p::CParams; device::Integer = 0, flags::UInt32 = FLAG_SPLIT_FFT)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:tfhe_b200_create, LIB), Cint, (Ref{CParams}, Cint, UInt32, Ref{Ptr{Cvoid}}), p, device, flags, h)
        rc == 0 || error("tfhe_b200_create: " * unsafe_string(ccall((:tfhe_b200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(h[], p)
        finalizer(c -> ccall((:tfhe_b200_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
        ctx
    end
end

function check(ctx::Context, rc::Cint)                          # error codes -> Julia exceptions (SURVEY §5)
    rc == 0 && return
    error("tfhe_b200 error $rc: " * unsafe_string(ccall((:tfhe_b200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.handle)))
end

cptr(a::Array{Int32}) = pointer(a)
cptr(::Nothing) = Ptr{Int32}(C_NULL)

# Julia arrays are column-major: Matrix{Int32}(n+1, count) IS the C layout [count][n+1].
function c_gate(ctx::Context, op::GateOp, x, y, z, count::Integer)
    out = Matrix{Int32}(undef, ctx.params.n * ctx.params.parties + 1, count)
    GC.@preserve x y z out check(ctx, ccall((:tfhe_b200_gate_batch, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Csize_t),
        ctx.handle, Int(op), cptr(x), cptr(y), cptr(z), out, count))
    out
end

function c_polymul(ctx::Context, x::Matrix{Int32}, y::Matrix{Int32})     # transformed_mul, polynomials.jl:142-144
    out = similar(x)
    GC.@preserve x y out check(ctx, ccall((:tfhe_b200_polymul_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Csize_t), ctx.handle, x, y, out, size(x, 2)))
    out
end

load_bk!(ctx::Context, bk::Array{Int32}) = GC.@preserve bk check(ctx, ccall(
    (ctx.params.parties == 1 ? :tfhe_b200_load_bk : :tfhe_b200_mk_load_bk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, bk))
load_ksk!(ctx::Context, ksk::Array{Int32}) = GC.@preserve ksk check(ctx, ccall(
    (ctx.params.parties == 1 ? :tfhe_b200_load_ksk : :tfhe_b200_mk_load_ksk, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ctx.handle, ksk))

# ---------------------------------------------------------------------------------------------- numeric-functions.jl
rand_uniform_bool(rng::AbstractRNG, dims...) = rand(rng, Int32(0):Int32(1), dims...)
rand_uniform_torus32(rng::AbstractRNG, dims...) = rand(rng, Torus32, dims...)
dtot32(d::Float64) = trunc(Int32, d * 2^32)
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

"An encrypted bit (lwe.jl:21-29) or a batch of them: `data` is (n+1) or (n+1, count); rows 1:n = a, row n+1 = b."
struct LweSample
    data::Array{Int32}
    current_variance::Float64
end
Base.length(x::LweSample) = size(x.data, 2)
Base.getindex(x::LweSample, i) = LweSample(x.data[:, i], x.current_variance)
Base.:+(x::LweSample, y::LweSample) = LweSample(x.data .+ y.data, x.current_variance + y.current_variance)   # lwe.jl:67-68
Base.:-(x::LweSample, y::LweSample) = LweSample(x.data .- y.data, x.current_variance + y.current_variance)   # lwe.jl:71-72
Base.:-(x::LweSample) = LweSample(.-x.data, x.current_variance)                                              # lwe.jl:74

struct SecretKey                                                # api.jl:92-100
    params::SchemeParameters
    key::Vector{Int32}
    SecretKey(rng::AbstractRNG, params::SchemeParameters) = new(params, rand_uniform_bool(rng, params.lwe_size))
end

# tlwe.jl:63-73 for `count` samples at once; the products S (*) a run on the GPU (kernel K1)
function tlwe_encrypt_zero(rng, ctx::Context, alpha::Float64, tlwe_key::Matrix{Int32}, count::Int)
    N, k = size(tlwe_key)
    a = rand_uniform_torus32(rng, N, k, count)
    b = rand_gaussian_torus32(rng, alpha, N, count)
    keys = repeat(reshape(tlwe_key, N, k, 1), 1, 1, count)
    prod = reshape(c_polymul(ctx, reshape(keys, N, k * count), reshape(a, N, k * count)), N, k, count)
    b .+= dropdims(sum(prod, dims = 2), dims = 2)
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
        ks[n + 1, h, j, i] = message + dtot32(noise[h, j, i]) + reduce(+, a .* out_key)                  # lwe.jl:49-55
    end
    ks
end

struct CloudKey                                                 # api.jl:111-127
    params::SchemeParameters
    ctx::Context
    bootstrap_key::Array{Int32}
    keyswitch_key::Array{Int32}
    function CloudKey(rng::AbstractRNG, secret_key::SecretKey; device::Integer = 0, flags::UInt32 = FLAG_SPLIT_FFT)
        p = secret_key.params
        ctx = Context(cparams(p, 1); device = device, flags = flags)
        tlwe_key = rand_uniform_bool(rng, p.tlwe_polynomial_degree, p.tlwe_mask_size)
        bk = bootstrap_key(rng, ctx, p.bs_noise_stddev, secret_key.key, tlwe_key, p.bs_decomp_length, p.bs_log2_base)
        ks = keyswitch_key(rng, p.ks_noise_stddev, p.ks_decomp_length, p.ks_log2_base, secret_key.key, vec(tlwe_key))
        load_bk!(ctx, bk); load_ksk!(ctx, ks)
        new(p, ctx, bk, ks)
    end
end

function make_key_pair(rng::AbstractRNG, params::Union{Nothing, SchemeParameters} = nothing; kwargs...)   # api.jl:139-146
    params === nothing && (params = tfhe_parameters_80())
    secret_key = SecretKey(rng, params)
    secret_key, CloudKey(rng, secret_key; kwargs...)
end

function lwe_encrypt(rng, message::Torus32, alpha::Float64, key::Vector{Int32})                           # lwe.jl:38-43
    a = rand_uniform_torus32(rng, length(key))
    vcat(a, message + dtot32(randn(rng) * alpha) + reduce(+, a .* key))
end

"api.jl:155-158; `message` may be a Bool or a vector of Bools (batched: one ciphertext per column)."
encrypt(rng::AbstractRNG, key::SecretKey, message::Bool) =
    LweSample(lwe_encrypt(rng, encode_message(message ? 1 : -1, 8), key.params.lwe_noise_stddev, key.key), key.params.lwe_noise_stddev^2)
encrypt(rng::AbstractRNG, key::SecretKey, messages::AbstractVector{Bool}) =
    LweSample(hcat([lwe_encrypt(rng, encode_message(m ? 1 : -1, 8), key.params.lwe_noise_stddev, key.key) for m in messages]...),
              key.params.lwe_noise_stddev^2)

lwe_phase(data::AbstractVector{Int32}, key) = data[end] - reduce(+, data[1:end-1] .* key)                 # lwe.jl:59
decrypt(key::SecretKey, sample::LweSample) = ndims(sample.data) == 1 ? lwe_phase(sample.data, key.key) > 0 :
    [lwe_phase(view(sample.data, :, g), key.key) > 0 for g in 1:size(sample.data, 2)]                     # api.jl:167-169

# ---------------------------------------------------------------------------------------------- gates.jl
as_batch(x::LweSample) = ndims(x.data) == 1 ? reshape(x.data, :, 1) : x.data
function gate(ck, op::GateOp, xs::LweSample...)
    mats = map(as_batch, xs)
    out = c_gate(ck.ctx, op, mats[1], length(mats) > 1 ? mats[2] : nothing, length(mats) > 2 ? mats[3] : nothing, size(mats[1], 2))
    LweSample(ndims(xs[1].data) == 1 ? vec(out) : out, 0.0)
end
gate_nand(ck, x, y) = gate(ck, NAND, x, y)          # gates.jl:15-18
gate_or(ck, x, y) = gate(ck, OR, x, y)              # gates.jl:27-30
gate_and(ck, x, y) = gate(ck, AND, x, y)            # gates.jl:39-42
gate_xor(ck, x, y) = gate(ck, XOR, x, y)            # gates.jl:51-54
gate_xnor(ck, x, y) = gate(ck, XNOR, x, y)          # gates.jl:63-66
gate_not(ck, x) = gate(ck, NOT, x)                  # gates.jl:76-79
gate_nor(ck, x, y) = gate(ck, NOR, x, y)            # gates.jl:102-105
gate_andny(ck, x, y) = gate(ck, ANDNY, x, y)        # gates.jl:114-117
gate_andyn(ck, x, y) = gate(ck, ANDYN, x, y)        # gates.jl:126-129
gate_orny(ck, x, y) = gate(ck, ORNY, x, y)          # gates.jl:138-141
gate_oryn(ck, x, y) = gate(ck, ORYN, x, y)          # gates.jl:150-153
gate_mux(ck, x, y, z) = gate(ck, MUX, x, y, z)      # gates.jl:163-177
function gate_constant(ck::CloudKey, value::Bool)   # gates.jl:91-93
    flags = zeros(Int32, ck.params.lwe_size + 1, 1); flags[1, 1] = value
    LweSample(vec(c_gate(ck.ctx, CONSTANT, flags, nothing, nothing, 1)), 0.0)
end

# ---------------------------------------------------------------------------------------------- multi-key
"mk_internals.jl:6-18: `data` is (p*n+1) or (p*n+1, count): a[:, party] blocks then the joint b."
struct MKLweSample
    data::Array{Int32}
    parties::Int
    current_variance::Float64
end

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
    offset = signed(UInt32(sum(Int64(1) << (32 - r * bgbit) for r in 1:l) * (1 << (bgbit - 1)) % 2^32))
    [((x .+ offset) .>> (32 - r * bgbit)) .& Int32((1 << bgbit) - 1) .- Int32(1 << (bgbit - 1)) for r in 1:l]
end

struct MKCloudKey                                               # mk_api.jl:85-101
    parties::Int
    params::SchemeParameters
    ctx::Context
    function MKCloudKey(ck_parts::Vector{CloudKeyPart}; flags::UInt32 = FLAG_SPLIT_FFT)
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
        ctx = Context(cparams(params, p); device = ck_parts[1].device, flags = flags)
        load_bk!(ctx, bk)
        load_ksk!(ctx, cat([part.ks for part in ck_parts]..., dims = 5))
        new(p, params, ctx)
    end
end

function mk_encrypt(rng, secret_keys::Vector{SecretKey}, message::Bool)                                   # mk_api.jl:110-126
    params = secret_keys[1].params
    keys = vcat([sk.key for sk in secret_keys]...)
    a = rand_uniform_torus32(rng, length(keys))
    b = encode_message(message ? 1 : -1, 8) + dtot32(randn(rng) * params.lwe_noise_stddev) + reduce(+, a .* keys)
    MKLweSample(vcat(a, b), length(secret_keys), params.lwe_noise_stddev^2)
end
mk_encrypt(rng, secret_keys::Vector{SecretKey}, messages::AbstractVector{Bool}) =
    MKLweSample(hcat([mk_encrypt(rng, secret_keys, m).data for m in messages]...), length(secret_keys), secret_keys[1].params.lwe_noise_stddev^2)

function mk_decrypt(secret_keys::Vector{SecretKey}, sample::MKLweSample)                                  # mk_api.jl:135-138
    keys = vcat([sk.key for sk in secret_keys]...)
    ndims(sample.data) == 1 ? lwe_phase(sample.data, keys) > 0 : [lwe_phase(view(sample.data, :, g), keys) > 0 for g in 1:size(sample.data, 2)]
end

function mk_gate_nand(ck::MKCloudKey, x::MKLweSample, y::MKLweSample)                                      # mk_gates.jl:7-12
    xm = ndims(x.data) == 1 ? reshape(x.data, :, 1) : x.data
    ym = ndims(y.data) == 1 ? reshape(y.data, :, 1) : y.data
    out = similar(xm)
    GC.@preserve xm ym out check(ck.ctx, ccall((:tfhe_b200_mk_nand_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Csize_t), ck.ctx.handle, xm, ym, out, size(xm, 2)))
    MKLweSample(ndims(x.data) == 1 ? vec(out) : out, ck.parties, 0.0)
end

end # module
