# DwaveHMCB200.jl -- thin `ccall` layer that forwards DwaveHMC.jl's operator API for the molecular-dynamics
# force path to libdwhmc.so (C ABI: include/dwhmc.h).
#
# HOW IT HOOKS IN.  This file is `include`d INSIDE `module DwaveHMC` (src/DwaveHMC.jl), after the package's own
# includes -- it is NOT a sub-module.  It therefore adds METHODS to the package's existing generic functions
# (init_static_H!, update_H_BdG!, diagonalize_H_BdG!, compute_forces!, compute_total_energy, hmc_sweep!,
# measure_observables, build_current_operator!, measure_transport_and_spectra), specialised on the new cache type
# `B200Cache`; the reference's methods are typed `cache::ComputeCache` (src/Hamiltonian.jl:10,55,96,
# src/Observables.jl:14,88,237,314, src/HMC.jl:12,71), so Julia's multiple dispatch selects by the cache argument
# and every call site of src/Simulation.jl and of the scan scripts stays as it is.  (Round 1 shipped this as a
# sub-module that exported functions of the same names: `using .DwaveHMCB200` inside DwaveHMC then kept the
# package's own bindings and `hmc_sweep!(::B200Cache, ...)` raised a MethodError.  Fixed here.)
#
# NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia (probed: `julia`, `juliaup` absent, no network).
# The file is a one-to-one transcription of hybrid-monte-carlo-for-d-wave-sc_b200/dwhmc/reference_api.py, which
# binds the same symbols through ctypes and is what the parity tests exercise
# (tests/test_gpu_parity.py::test_golden_vectors_single_chain).
#
# Reference-side change (see INTEGRATION.md section 2):
#     # src/DwaveHMC.jl, after include("Simulation.jl")
#     include("DwaveHMCB200.jl")
#     export B200Cache, B200Batch, fetch_eigensystem!
#     # src/Simulation.jl:82      cache = B200Cache(p)       (was: initialize_cache(p))
using Random

const DWHMC_LIB = get(ENV, "DWHMC_LIB", joinpath(@__DIR__, "..", "libdwhmc.so"))

struct DwhmcError <: Exception
    code::Cint
    msg::String
end

function dwhmc_check(h::Ptr{Cvoid}, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:dwhmc_last_error, DWHMC_LIB), Cstring, (Ptr{Cvoid},), h))
    throw(DwhmcError(rc, msg))          # the reference's error convention is exceptions (LAPACKException at :106)
end

"""One chain on one GPU (the drop-in for `ComputeCache`).  Holds the device handle; `E_n`, `U`, `forces`,
`fermi_factors` are host mirrors (same field names as `ComputeCache`, src/Types.jl:145-180) filled by
`compute_forces!` (forces) and `fetch_eigensystem!` (the rest; nothing on the hot path needs them on the host).
`dirty_Δ` / `dirty_π` avoid re-sending the 18 KB field to the device when the caller has not touched it."""
mutable struct B200Cache
    h::Ptr{Cvoid}
    N::Int
    E_n::Vector{Float64}
    U::Matrix{ComplexF64}
    forces::Matrix{ComplexF64}
    fermi_factors::Vector{Float64}
    params::NTuple{6,Float64}
    Δ_sent::Matrix{ComplexF64}      # what the device holds (to skip redundant host -> device copies)
    have_Δ::Bool
end

function B200Cache(p::ModelParameters; device::Integer=0)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    nn = Matrix{Int64}(p.nn_table); nnn = Matrix{Int64}(p.nnn_table)      # N x 4, column-major, 1-based
    rc = ccall((:dwhmc_create, DWHMC_LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Ptr{Int64}, Ptr{Int64}),
               href, device, 1, p.Lx, p.Ly, nn, nnn)
    dwhmc_check(Ptr{Cvoid}(C_NULL), rc)
    dim = 2 * p.N
    c = B200Cache(href[], p.N, zeros(dim), zeros(ComplexF64, dim, dim), zeros(ComplexF64, p.N, 2), zeros(dim),
                  (NaN, NaN, NaN, NaN, NaN, NaN), zeros(ComplexF64, p.N, 2), false)
    finalizer(x -> ccall((:dwhmc_destroy, DWHMC_LIB), Cint, (Ptr{Cvoid},), x.h), c)
    return c
end

function sync_params!(c::B200Cache, p::ModelParameters)
    key = (Float64(p.t), Float64(p.tp), Float64(p.μ), Float64(p.β), Float64(p.J), Float64(p.mass))
    if key != c.params
        dwhmc_check(c.h, ccall((:dwhmc_set_params, DWHMC_LIB), Cint,
                               (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}),
                               c.h, key[1], key[2], key[3], key[4], key[5], key[6]))
        c.params = key
    end
end

function push_field!(c::B200Cache, state::SimulationState)
    if !c.have_Δ || c.Δ_sent != state.Δ             # 2N complex numbers: the comparison is cheaper than the copy
        dwhmc_check(c.h, ccall((:dwhmc_set_field, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, state.Δ))
        copyto!(c.Δ_sent, state.Δ); c.have_Δ = true
    end
end

# init_static_H!  (src/Hamiltonian.jl:10-47)
function init_static_H!(c::B200Cache, p::ModelParameters, state::SimulationState)
    sync_params!(c, p)
    dwhmc_check(c.h, ccall((:dwhmc_set_disorder, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, state.disorder_pot))
    dwhmc_check(c.h, ccall((:dwhmc_init_static_H, DWHMC_LIB), Cint, (Ptr{Cvoid},), c.h))
    return nothing
end

# update_H_BdG!  (src/Hamiltonian.jl:55-86)
function update_H_BdG!(c::B200Cache, p::ModelParameters, state::SimulationState)
    push_field!(c, state)
    dwhmc_check(c.h, ccall((:dwhmc_update_H, DWHMC_LIB), Cint, (Ptr{Cvoid},), c.h))
    return nothing
end

# diagonalize_H_BdG!  (src/Hamiltonian.jl:96-114)
function diagonalize_H_BdG!(c::B200Cache, p::ModelParameters)
    dwhmc_check(c.h, ccall((:dwhmc_diagonalize, DWHMC_LIB), Cint, (Ptr{Cvoid},), c.h))
    return nothing
end

# compute_forces!  (src/Observables.jl:14-62)
function compute_forces!(c::B200Cache, p::ModelParameters, state::SimulationState)
    sync_params!(c, p); push_field!(c, state)
    dwhmc_check(c.h, ccall((:dwhmc_compute_forces, DWHMC_LIB), Cint, (Ptr{Cvoid},), c.h))
    dwhmc_check(c.h, ccall((:dwhmc_get_forces, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, c.forces))
    return nothing
end

# compute_total_energy  (src/HMC.jl:12-41)
function compute_total_energy(c::B200Cache, p::ModelParameters, state::SimulationState)
    sync_params!(c, p); push_field!(c, state)
    dwhmc_check(c.h, ccall((:dwhmc_set_momentum, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, state.π))
    out = Ref{Float64}(0.0)
    dwhmc_check(c.h, ccall((:dwhmc_total_energy, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ref{Float64}), c.h, out))
    return out[]
end

# hmc_sweep!  (src/HMC.jl:71-144): same RNG consumption as the reference -- randn!(π) first, then rand() only when
# ΔH >= 0 (:53, :128) -- because trajectory and commit are separate C calls.
function hmc_sweep!(c::B200Cache, p::ModelParameters, state::SimulationState; Nt::Int, dt::Float64)
    sync_params!(c, p)
    randn!(state.π)
    state.π .*= sqrt(2 * p.mass)
    push_field!(c, state)
    dH = Ref{Float64}(0.0)
    dwhmc_check(c.h, ccall((:dwhmc_trajectory, DWHMC_LIB), Cint,
                           (Ptr{Cvoid}, Ref{Int32}, Ref{Float64}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                           c.h, Int32(Nt), dt, state.π, C_NULL, C_NULL, dH))
    ΔH = dH[]
    accepted = (ΔH < 0 || rand() < exp(-ΔH))
    dwhmc_check(c.h, ccall((:dwhmc_commit, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ref{Int32}), c.h, Int32(accepted)))
    dwhmc_check(c.h, ccall((:dwhmc_get_field, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, state.Δ))
    dwhmc_check(c.h, ccall((:dwhmc_get_momentum, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, state.π))
    copyto!(c.Δ_sent, state.Δ); c.have_Δ = true      # the device already holds this field
    return accepted, ΔH
end

# measure_observables  (src/Observables.jl:88-222)
function measure_observables(c::B200Cache, p::ModelParameters, state::SimulationState)
    sync_params!(c, p); push_field!(c, state)
    out = zeros(9)
    dwhmc_check(c.h, ccall((:dwhmc_measure_observables, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, out))
    return ObservablesResult(out...)                 # field order of src/Observables.jl:70-80
end

# build_current_operator!  (src/Observables.jl:237-283): the operator lives in the kernels
build_current_operator!(c::B200Cache, p::ModelParameters) = (sync_params!(c, p); nothing)

# measure_transport_and_spectra  (src/Observables.jl:314-526)
function measure_transport_and_spectra(c::B200Cache, p::ModelParameters)
    sync_params!(c, p)
    ω = collect(p.ω_min:p.Δω:p.ω_max); ωd = collect(-p.ω_max:p.Δω:p.ω_max)      # :396, :432
    scal = zeros(2); σ = zeros(length(ω)); dos = zeros(length(ωd)); dosAN = zeros(length(ωd))
    Ak0 = zeros(p.Lx, p.Ly)
    dwhmc_check(c.h, ccall((:dwhmc_measure_transport, DWHMC_LIB), Cint,
                           (Ptr{Cvoid}, Cdouble, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64},
                            Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                           c.h, p.η, ω, length(ω), ωd, length(ωd), scal, σ, dos, dosAN, Ak0))
    return SpectrumResult(scal[1], scal[2], ω, σ, ωd, dos, dosAN, Ak0)         # field order of :293-311
end

"""Copy E_n, U and the Fermi factors back to the host mirrors (debugging; the hot path never needs them on the host)."""
function fetch_eigensystem!(c::B200Cache)
    dwhmc_check(c.h, ccall((:dwhmc_get_eigenvalues, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, c.E_n))
    dwhmc_check(c.h, ccall((:dwhmc_get_eigenvectors, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, c.U))
    dwhmc_check(c.h, ccall((:dwhmc_get_fermi, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, c.fermi_factors))
    return nothing
end

# ---------------------------------------------------------------------------------------------------------------
# B chains in one handle: the primitive the batched scan driver (julia/batch_scan_T_b200.jl) is built on.
# Array layouts: chain index last (Julia column-major = "chain slowest" of include/dwhmc.h).
# ---------------------------------------------------------------------------------------------------------------
mutable struct B200Batch
    h::Ptr{Cvoid}
    B::Int
    N::Int
end

function B200Batch(p::ModelParameters, B::Integer; device::Integer=0)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    nn = Matrix{Int64}(p.nn_table); nnn = Matrix{Int64}(p.nnn_table)
    rc = ccall((:dwhmc_create, DWHMC_LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Ptr{Int64}, Ptr{Int64}),
               href, device, B, p.Lx, p.Ly, nn, nnn)
    dwhmc_check(Ptr{Cvoid}(C_NULL), rc)
    b = B200Batch(href[], B, p.N)
    finalizer(x -> ccall((:dwhmc_destroy, DWHMC_LIB), Cint, (Ptr{Cvoid},), x.h), b)
    return b
end

batch_set_params!(b::B200Batch, t::Vector{Float64}, tp::Vector{Float64}, μ::Vector{Float64}, β::Vector{Float64},
                  J::Vector{Float64}, mass::Vector{Float64}) =
    dwhmc_check(b.h, ccall((:dwhmc_set_params, DWHMC_LIB), Cint,
                           (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                           b.h, t, tp, μ, β, J, mass))
batch_seed!(b::B200Batch, seed::Integer) = dwhmc_check(b.h, ccall((:dwhmc_seed, DWHMC_LIB), Cint, (Ptr{Cvoid}, UInt64), b.h, seed))
# disorder [N, B], field / momentum [N, 2, B]
batch_set_disorder!(b::B200Batch, w::Matrix{Float64}) =
    dwhmc_check(b.h, ccall((:dwhmc_set_disorder, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.h, w))
batch_set_field!(b::B200Batch, Δ::Array{ComplexF64,3}) =
    dwhmc_check(b.h, ccall((:dwhmc_set_field, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), b.h, Δ))
function batch_prepare!(b::B200Batch)               # init_static_H!, update_H_BdG!, diagonalize_H_BdG! for every chain
    dwhmc_check(b.h, ccall((:dwhmc_init_static_H, DWHMC_LIB), Cint, (Ptr{Cvoid},), b.h))
    dwhmc_check(b.h, ccall((:dwhmc_update_H, DWHMC_LIB), Cint, (Ptr{Cvoid},), b.h))
    dwhmc_check(b.h, ccall((:dwhmc_diagonalize, DWHMC_LIB), Cint, (Ptr{Cvoid},), b.h))
end
"""hmc_sweep! for every chain with the on-device RNG (momenta and Metropolis uniforms); per-chain Nt and dt."""
function batch_hmc_sweep!(b::B200Batch, Nt::Vector{Int32}, dt::Vector{Float64})
    acc = zeros(Int32, b.B); dH = zeros(b.B)
    dwhmc_check(b.h, ccall((:dwhmc_hmc_sweep, DWHMC_LIB), Cint,
                           (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}),
                           b.h, Nt, dt, C_NULL, C_NULL, acc, dH))
    return acc .!= 0, dH
end
function batch_measure_observables(b::B200Batch)     # [9, B], rows in ObservablesResult field order
    out = zeros(9, b.B)
    dwhmc_check(b.h, ccall((:dwhmc_measure_observables, DWHMC_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.h, out))
    return out
end
function batch_measure_transport(b::B200Batch, p::ModelParameters)
    ω = collect(p.ω_min:p.Δω:p.ω_max); ωd = collect(-p.ω_max:p.Δω:p.ω_max)
    scal = zeros(2, b.B); σ = zeros(length(ω), b.B); dos = zeros(length(ωd), b.B); dosAN = zeros(length(ωd), b.B)
    Ak0 = zeros(p.Lx, p.Ly, b.B)
    dwhmc_check(b.h, ccall((:dwhmc_measure_transport, DWHMC_LIB), Cint,
                           (Ptr{Cvoid}, Cdouble, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64},
                            Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                           b.h, p.η, ω, length(ω), ωd, length(ωd), scal, σ, dos, dosAN, Ak0))
    return [SpectrumResult(scal[1, c], scal[2, c], ω, σ[:, c], ωd, dos[:, c], dosAN[:, c], Ak0[:, :, c]) for c in 1:b.B]
end
