# DwaveHMCB200.jl -- thin `ccall` shim that keeps DwaveHMC.jl's operator API for the molecular-
# dynamics force path and forwards it to libdwhmc.so (C ABI: include/dwhmc.h).
#
# UNTESTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The file is deliberately a
# one-to-one transcription of hybrid-monte-carlo-for-d-wave-sc_b200/dwhmc/reference_api.py, which
# binds the same symbols through ctypes and is what the parity tests exercise.
#
# Usage inside the reference package (see INTEGRATION.md):
#     include("DwaveHMCB200.jl"); using .DwaveHMCB200
#     cache = B200Cache(p)                      # instead of initialize_cache(p)
#     init_static_H!(cache, p, state); update_H_BdG!(cache, p, state); diagonalize_H_BdG!(cache, p)
#     accepted, dH = hmc_sweep!(cache, p, state; Nt=6, dt=dt)
# `p::ModelParameters` and `state::SimulationState` are the reference's own structs (src/Types.jl),
# unchanged; only the cache type differs, so Julia's dispatch picks these methods.
module DwaveHMCB200

using Random

export B200Cache, init_static_H!, update_H_BdG!, diagonalize_H_BdG!, compute_forces!,
       compute_total_energy, hmc_sweep!, measure_observables, fetch_eigensystem!

const LIB = get(ENV, "DWHMC_LIB", joinpath(@__DIR__, "..", "libdwhmc.so"))

struct DwhmcError <: Exception
    code::Cint
    msg::String
end

function check(h::Ptr{Cvoid}, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:dwhmc_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    throw(DwhmcError(rc, msg))          # the reference's error convention is exceptions
end

"""One chain on one GPU.  Holds the device handle; `E_n`, `U`, `forces`, `fermi_factors` are host
mirrors filled on demand by `fetch_eigensystem!`."""
mutable struct B200Cache
    h::Ptr{Cvoid}
    N::Int
    E_n::Vector{Float64}
    U::Matrix{ComplexF64}
    forces::Matrix{ComplexF64}
    fermi_factors::Vector{Float64}
    params::NTuple{6,Float64}
end

function B200Cache(p; device::Integer=0)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    nn = Matrix{Int64}(p.nn_table); nnn = Matrix{Int64}(p.nnn_table)      # N x 4, column-major, 1-based
    rc = ccall((:dwhmc_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Ptr{Int64}, Ptr{Int64}),
               href, device, 1, p.Lx, p.Ly, nn, nnn)
    check(Ptr{Cvoid}(C_NULL), rc)
    dim = 2 * p.N
    c = B200Cache(href[], p.N, zeros(dim), zeros(ComplexF64, dim, dim), zeros(ComplexF64, p.N, 2), zeros(dim),
                  (NaN, NaN, NaN, NaN, NaN, NaN))
    finalizer(x -> ccall((:dwhmc_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
    return c
end

function sync_params!(c::B200Cache, p)
    key = (p.t, p.tp, p.μ, p.β, p.J, p.mass)
    if key != c.params
        check(c.h, ccall((:dwhmc_set_params, LIB), Cint,
                         (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}),
                         c.h, p.t, p.tp, p.μ, p.β, p.J, p.mass))
        c.params = key
    end
end

push_field!(c::B200Cache, state) =
    check(c.h, ccall((:dwhmc_set_field, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, state.Δ))

# init_static_H!  (src/Hamiltonian.jl:10-47)
function init_static_H!(c::B200Cache, p, state)
    sync_params!(c, p)
    check(c.h, ccall((:dwhmc_set_disorder, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, state.disorder_pot))
    check(c.h, ccall((:dwhmc_init_static_H, LIB), Cint, (Ptr{Cvoid},), c.h))
    return nothing
end

# update_H_BdG!  (src/Hamiltonian.jl:55-86)
function update_H_BdG!(c::B200Cache, p, state)
    push_field!(c, state)
    check(c.h, ccall((:dwhmc_update_H, LIB), Cint, (Ptr{Cvoid},), c.h))
    return nothing
end

# diagonalize_H_BdG!  (src/Hamiltonian.jl:96-114)
function diagonalize_H_BdG!(c::B200Cache, p)
    check(c.h, ccall((:dwhmc_diagonalize, LIB), Cint, (Ptr{Cvoid},), c.h))
    return nothing
end

# compute_forces!  (src/Observables.jl:14-62)
function compute_forces!(c::B200Cache, p, state)
    sync_params!(c, p); push_field!(c, state)
    check(c.h, ccall((:dwhmc_compute_forces, LIB), Cint, (Ptr{Cvoid},), c.h))
    check(c.h, ccall((:dwhmc_get_forces, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, c.forces))
    return nothing
end

# compute_total_energy  (src/HMC.jl:12-41)
function compute_total_energy(c::B200Cache, p, state)
    sync_params!(c, p); push_field!(c, state)
    check(c.h, ccall((:dwhmc_set_momentum, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, state.π))
    out = Ref{Float64}(0.0)
    check(c.h, ccall((:dwhmc_total_energy, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}), c.h, out))
    return out[]
end

# hmc_sweep!  (src/HMC.jl:71-144): same RNG consumption as the reference -- randn!(π) first, then
# rand() only when ΔH >= 0 -- because trajectory and commit are separate C calls.
function hmc_sweep!(c::B200Cache, p, state; Nt::Int, dt::Float64)
    sync_params!(c, p)
    randn!(state.π)
    state.π .*= sqrt(2 * p.mass)
    push_field!(c, state)
    dH = Ref{Float64}(0.0)
    check(c.h, ccall((:dwhmc_trajectory, LIB), Cint,
                     (Ptr{Cvoid}, Ref{Int32}, Ref{Float64}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                     c.h, Int32(Nt), dt, state.π, C_NULL, C_NULL, dH))
    ΔH = dH[]
    accepted = (ΔH < 0 || rand() < exp(-ΔH))
    check(c.h, ccall((:dwhmc_commit, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}), c.h, Int32(accepted)))
    check(c.h, ccall((:dwhmc_get_field, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, state.Δ))
    check(c.h, ccall((:dwhmc_get_momentum, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, state.π))
    return accepted, ΔH
end

# measure_observables  (src/Observables.jl:88-222); returns the 9 fields in ObservablesResult order
function measure_observables(c::B200Cache, p, state)
    sync_params!(c, p); push_field!(c, state)
    out = zeros(9)
    check(c.h, ccall((:dwhmc_measure_observables, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, out))
    return out        # wrap as DwaveHMC.ObservablesResult(out...) at the call site
end

# measure_transport_and_spectra + build_current_operator!  (src/Observables.jl:314-526, :237-283).
# Returns the fields of SpectrumResult (:293-311) in order; wrap as DwaveHMC.SpectrumResult(r...) at the call site.
build_current_operator!(c::B200Cache, p) = (sync_params!(c, p); nothing)   # the operator lives in the kernels
function measure_transport_and_spectra(c::B200Cache, p)
    sync_params!(c, p)
    ω = collect(p.ω_min:p.Δω:p.ω_max); ωd = collect(-p.ω_max:p.Δω:p.ω_max)
    scal = zeros(2); σ = zeros(length(ω)); dos = zeros(length(ωd)); dosAN = zeros(length(ωd))
    Ak0 = zeros(p.Lx, p.Ly)
    check(c.h, ccall((:dwhmc_measure_transport, LIB), Cint,
                     (Ptr{Cvoid}, Cdouble, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64},
                      Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     c.h, p.η, ω, length(ω), ωd, length(ωd), scal, σ, dos, dosAN, Ak0))
    return (scal[1], scal[2], ω, σ, ωd, dos, dosAN, Ak0)
end

"""Copy E_n and U back to the host mirrors (debugging; the transport path no longer needs them on the host)."""
function fetch_eigensystem!(c::B200Cache)
    check(c.h, ccall((:dwhmc_get_eigenvalues, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, c.E_n))
    check(c.h, ccall((:dwhmc_get_eigenvectors, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), c.h, c.U))
    check(c.h, ccall((:dwhmc_get_fermi, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, c.fermi_factors))
    return nothing
end

end # module
