# SphB200.jl -- the `ccall` shim a maintainer of george-toka/Astrophysical-SPH adds next to sph_manager.jl to run the
# per-step core on a B200 through libsph_b200.so (include/sph_b200.h).  NOT executed in this repository's CI: the build
# image has no Julia toolchain (DESIGN.md section 2); the Python host in astrophysical-sph_b200/ mirrors it call for call.
#
# Usage inside isothermal_sim.jl / polytrope_sim.jl (julia_version/fastv1_kd&single_oc):
#     include("SphB200.jl"); using .SphB200
#     h = SphB200.create(N, Kh, :isothermal; m, cs, G, theta, alpha, beta, U)      # replaces the unpacking at :87-105
#     SphB200.upload!(h, Matrix{Float64}(pos), Matrix{Float64}(vel), nothing, t)  # convert: read_snapshot may return
#                                                                                 # Matrix{Union{Missing,Float64}}
#     while t < tEnd
#         dt, stats_vector = SphB200.step!(h)                                     # replaces lines :155-212
#         SnapshotRW.update_stats_row!(stats_arr, iterID, stats_vector); t += dt
#         ...snapshot logic unchanged; pos, vel = SphB200.download(h) when a snapshot is due...
#     end
module SphB200

const LIB = get(ENV, "SPH_B200_LIB", "libsph_b200.so")

struct Params          # mirrors `sph_params` (include/sph_b200.h), 88 bytes, C layout
    N::Int64
    Kh::Int32
    eos::Int32
    m::Float64
    cs::Float64
    gamma::Float64
    G::Float64
    theta::Float64
    alpha::Float64
    beta::Float64
    U_iso::Float64
    device::Int32
    flags::Int32       # FLAG_* bits
end
const FLAG_COUNT_VISITS = Int32(1)    # count the node visits of the tree walk (timings: walk_visits)
const FLAG_SERIAL_PHASES = Int32(2)   # density / force before the walk on one stream: phase timers of every kernel alone

struct StepInfo        # mirrors `sph_step_info`
    dt::Float64
    stats::NTuple{10,Float64}
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    N::Int
    poly::Bool
end

function check(h, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:sph_last_error, LIB), Cstring, (Ptr{Cvoid},), h === nothing ? C_NULL : h.ptr))
    error("libsph_b200 error $rc: $msg")
end

const ABI_VERSION = 2      # SPH_B200_ABI_VERSION of include/sph_b200.h this file was written against
const ERR_NAN = -8         # dt came out NaN: the reference's loop ends here (minimum() propagates NaN, `while t < tEnd` fails)

function create(N::Integer, Kh::Integer, eos::Symbol; m, cs=0.0, gamma=5/3, G, theta, alpha, beta, U=0.0, device=0)
    v = ccall((:sph_abi_version, LIB), Cint, ())
    v == ABI_VERSION || error("libsph_b200 ABI version $v, expected $ABI_VERSION")
    p = Ref(Params(N, Kh, eos == :polytropic ? 1 : 0, m, cs, gamma, G, theta, alpha, beta, U, device, 0))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:sph_create, LIB), Cint, (Ref{Params}, Ref{Ptr{Cvoid}}), p, out)
    check(nothing, rc)
    h = Handle(out[], N, eos == :polytropic)
    finalizer(x -> ccall((:sph_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

# pos, vel: Matrix{Float64} N x 3 (column-major = the ABI's layout, no copy); K: Vector{Float64} or nothing
function upload!(h::Handle, pos::Matrix{Float64}, vel::Matrix{Float64}, K, t::Float64)
    Kp = K === nothing ? Ptr{Float64}(C_NULL) : pointer(K)
    GC.@preserve pos vel K check(h, ccall((:sph_upload, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble), h.ptr, pos, vel, Kp, t))
end

function download(h::Handle)
    pos = Matrix{Float64}(undef, h.N, 3); vel = similar(pos)
    K = h.poly ? Vector{Float64}(undef, h.N) : nothing
    t = Ref{Cdouble}(0.0)
    Kp = K === nothing ? Ptr{Float64}(C_NULL) : pointer(K)
    GC.@preserve pos vel K check(h, ccall((:sph_download, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Cdouble}), h.ptr, pos, vel, Kp, t))
    return pos, vel, K, t[]
end

# one iteration of `while t < tEnd` (isothermal_sim.jl:155-212): returns (dt, stats row)
function step!(h::Handle)
    info = Ref(StepInfo(0.0, ntuple(_ -> 0.0, 10)))
    check(h, ccall((:sph_step, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{StepInfo}), h.ptr, 1, info))
    return info[].dt, collect(info[].stats)
end

# drop-in for getAcc (isothermal_sim.jl:16-49): acc, rho, h, PHI for caller-supplied pos/vel
function getAcc(h::Handle, pos::Matrix{Float64}, vel::Matrix{Float64}, K=nothing)
    acc = Matrix{Float64}(undef, h.N, 3); rho = Vector{Float64}(undef, h.N); hs = similar(rho); phi = similar(rho)
    Kp = K === nothing ? Ptr{Float64}(C_NULL) : pointer(K)
    GC.@preserve pos vel K check(h, ccall((:sph_eval_acc, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        h.ptr, pos, vel, Kp, acc, rho, hs, phi))
    return acc, rho, hs, phi
end

# HJL.density_plot (isothermal_hydroKDTree.jl:291-297) on the uploaded positions
function density_plot(h::Handle, rr::Matrix{Float64})
    out = Vector{Float64}(undef, size(rr, 1))
    check(h, ccall((:sph_density_at, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), h.ptr, rr, size(rr, 1), out))
    return out
end

# HJL.getNeighbors' index matrix (isothermal_hydroKDTree.jl:136-142): N x Kh Int32, 1-based, column 1 = self
function neighbors(h::Handle, Kh::Integer)
    idx = Matrix{Int32}(undef, h.N, Kh); r = Matrix{Float64}(undef, h.N, Kh)
    check(h, ccall((:sph_get_neighbors, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}), h.ptr, idx, r))
    return idx, r
end

end # module
