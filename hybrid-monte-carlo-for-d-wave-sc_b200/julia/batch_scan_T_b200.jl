# batch_scan_T_b200.jl -- scripts/batch_scan_T.jl with every temperature point (x disorder seeds) of the scan
# advancing at once as one B-chain handle on the GPU.
#
# The reference loops `for (i, T) in enumerate(Ts) ... run_simulation(p, work_dir; ...) end`
# (scripts/batch_scan_T.jl:54-74), one chain after the other.  The chains are independent Markov chains, so here
# chain c = (i_T - 1) * n_seeds + i_seed lives in slot c of a `B200Batch` (julia/DwaveHMCB200.jl): per-chain β,
# disorder, field, leapfrog step count Nt and step dt.  What `run_simulation` (src/Simulation.jl:34-236) does per
# chain is kept per chain: adaptive thermalisation (window of 5 sweeps: rate < 0.60 -> Nt += 2; rate > 0.95 and
# Nt > 4 -> Nt -= 1; dt = calc_optimal_dt(β, J, m, Nt), :109-124), the measurement loop at Nt_measure, one directory per
# chain (`T_$(round(T, sigdigits=3))`, with more than one seed `.../seed_$k`) holding simulation.log,
# observables.csv, transport.csv and spectra_bins.jld2 in the reference's formats (:71-73, :161-165, :174-175,
# :89, :206-214), so scripts/batch_csv_summary_T.jl and scripts/*process_spectra.jl read them unchanged.
# Momenta and Metropolis uniforms come from the device RNG (Philox, seeded per handle); the initial state is drawn
# on the host with the reference's own `initialize_state(p)`.
#
# NOT EXECUTED IN THIS REPOSITORY (no Julia in the build image).  Its executed twin is
# hybrid-monte-carlo-for-d-wave-sc_b200/dwhmc/simulation.py (`run_simulation_batch`, `batch_scan_T`), tested against
# a sequential replay of the oracle in tests/test_gpu_parity.py::test_batched_run_driver_matches_oracle_run.
#
# Multi-GPU: start one Julia process per GPU with DWHMC_RANK / DWHMC_WORLD set; rank r takes chains r+1, r+1+W, ...
# (round robin, as dwhmc/parallel.py does); nothing is exchanged during the run, every chain writes its own files.
using DwaveHMC
using Printf
using Dates
using JLD2

# ---- scan definition: scripts/batch_scan_T.jl:10-36 -----------------------------------------------------------
Lx, Ly = 24, 24
t, tp = 1.0, -0.35
μ = -1.08
W, n_imp = 1.0, 0.05
J = 0.8
mass = 1.0
η = 8.0 / (Lx * Ly) * 1.0
Δω = 0.2 * η
ω_max = 4.0
T_start, T_end, n_points = 0.0001, 1000.0, 24
Ts = 10 .^ range(log10(T_start), stop=log10(T_end), length=n_points)
n_seeds = parse(Int, get(ENV, "DWHMC_SEEDS", "1"))
n_therm, n_measure = 20, 100
Nt_therm, Nt_measure = 20, 6
measure_freq, bin_size = 1, 10
base_dir = "data/T_scan_L$(Lx)_J$(J)_W$(W)_imp$(n_imp)_mu_$(μ)"

rank = parse(Int, get(ENV, "DWHMC_RANK", "0")); world = parse(Int, get(ENV, "DWHMC_WORLD", "1"))
all_chains = [(iT, k) for iT in 1:n_points for k in 1:n_seeds]
mine = all_chains[(rank + 1):world:end]
B = length(mine)

# ---- per-chain parameters, states, files ------------------------------------------------------------------------
ps = [ModelParameters(Lx, Ly, t, tp, μ, W, n_imp, 1.0 / Ts[iT], J, mass, η=η, Δω=Δω, ω_max=ω_max) for (iT, _) in mine]
dirs = [n_seeds == 1 ? joinpath(base_dir, "T_$(round(Ts[iT], sigdigits=3))") :
                       joinpath(base_dir, "T_$(round(Ts[iT], sigdigits=3))", "seed_$k") for (iT, k) in mine]
states = [initialize_state(p) for p in ps]          # disorder sites and Δ0 as the reference draws them (src/Types.jl:118-134)
N = ps[1].N
foreach(mkpath, dirs)
f_log = [open(joinpath(d, "simulation.log"), "a") for d in dirs]
f_obs = [open(joinpath(d, "observables.csv"), "w") for d in dirs]
f_trans = [open(joinpath(d, "transport.csv"), "w") for d in dirs]
jld = [joinpath(d, "spectra_bins.jld2") for d in dirs]
function tee(c, msg)
    println(f_log[c], "[$(Dates.format(now(), "yyyy-mm-dd HH:MM:SS"))] $msg"); flush(f_log[c])
end
for c in 1:B
    println(f_obs[c], "Sweep,Accepted,dH,Energy,Delta_Amp,Delta_Loc,Delta_Glob,S_Delta,Hole_p,Delta_Diff,Delta_Pair,Delta_LocalPair")
    println(f_trans[c], "Sweep,Superfluid_Stiffness,DC_Conductivity")
    tee(c, "Starting Simulation...")
    tee(c, "System: $(Lx)x$(Ly), β=$(ps[c].β), n_imp=$(n_imp), J=$(J)")
    tee(c, "Config: Therm=$n_therm, Sweep=$n_measure, TransFreq=$measure_freq, BinSize=$bin_size")
    jldsave(jld[c]; params=ps[c], omega_grid=collect(ps[c].ω_min:ps[c].Δω:ps[c].ω_max))
end

# ---- one handle for all chains of this rank -----------------------------------------------------------------------
b = DwaveHMC.B200Batch(ps[1], B; device=parse(Int, get(ENV, "DWHMC_DEVICE", "0")))
DwaveHMC.batch_set_params!(b, fill(t, B), fill(tp, B), fill(μ, B), [p.β for p in ps], fill(J, B), fill(mass, B))
wmat = zeros(N, B); Δ0 = zeros(ComplexF64, N, 2, B)
for c in 1:B
    wmat[:, c] .= states[c].disorder_pot
    Δ0[:, :, c] .= states[c].Δ
end
DwaveHMC.batch_set_disorder!(b, wmat); DwaveHMC.batch_set_field!(b, Δ0)
DwaveHMC.batch_prepare!(b)                          # init_static_H!, update_H_BdG!, diagonalize_H_BdG!  (:84-86)
DwaveHMC.batch_seed!(b, 0x5eed0000 + rank)

# ---- adaptive thermalisation, per chain (src/Simulation.jl:92-130) ---------------------------------------------------
Nt = fill(Int32(Nt_therm), B)
dt = [calc_optimal_dt(p.β, p.J, p.mass, Nt_therm) for p in ps]
for c in 1:B
    tee(c, "--- Thermalization Start ---"); tee(c, "Init: Nt=$(Nt[c]), dt=$(round(dt[c], digits=5))")
end
therm_window = 5
recent = zeros(Int, B)
for i in 1:n_therm
    acc, _ = DwaveHMC.batch_hmc_sweep!(b, Nt, dt)
    recent .+= acc
    if i % therm_window == 0
        for c in 1:B
            rate = recent[c] / therm_window
            old = Nt[c]
            if rate < 0.60
                Nt[c] += 2
            elseif rate > 0.95 && Nt[c] > 4
                Nt[c] -= 1
            end
            if Nt[c] != old
                dt[c] = calc_optimal_dt(ps[c].β, ps[c].J, ps[c].mass, Nt[c])
                tee(c, @sprintf("Therm %d/%d. Rate=%.2f. Adjust Nt: %d -> %d, dt: %.4f", i, n_therm, rate, old, Nt[c], dt[c]))
            elseif i % 20 == 0
                tee(c, @sprintf("Therm %d/%d. Rate=%.2f. Nt=%d (Stable)", i, n_therm, rate, Nt[c]))
            end
        end
        recent .= 0
    end
end

# ---- measurement (src/Simulation.jl:134-228) --------------------------------------------------------------------------
Nt .= Int32(Nt_measure)
dt .= [calc_optimal_dt(p.β, p.J, p.mass, Nt_measure) for p in ps]
for c in 1:B
    tee(c, "--- Measurement Start ---"); tee(c, "Settings: Nt=$Nt_measure, dt=$(round(dt[c], digits=5))")
end
acc_total = zeros(Int, B)
bins = [Dict{Symbol,Any}(:count => 0) for _ in 1:B]
for i in 1:n_measure
    acc, dH = DwaveHMC.batch_hmc_sweep!(b, Nt, dt)
    acc_total .+= acc
    obs = DwaveHMC.batch_measure_observables(b)     # [9, B], ObservablesResult field order: total_energy, Δ_amp, Δ_local, Δ_global, S_Δ, hole_conc, Δ_diff, Δ_pair, Δ_localpair
    for c in 1:B
        o = ObservablesResult(obs[:, c]...)
        write(f_obs[c], @sprintf("%d,%d,%.5e,%.6f,%.6f,%.6f,%.6f,%.6f,%.6f,%.6f,%.6f,%.6f\n", i, acc[c], dH[c],
                                 o.total_energy, o.Δ_amp, o.Δ_local, o.Δ_global, o.S_Δ, o.hole_conc, o.Δ_diff, o.Δ_pair,
                                 o.Δ_localpair))
        flush(f_obs[c])
    end
    if i % measure_freq == 0
        specs = DwaveHMC.batch_measure_transport(b, ps[1])      # one batched call: J = U†(Jx U) on the tensor cores
        for c in 1:B
            s = specs[c]
            write(f_trans[c], @sprintf("%d,%.6f,%.6f\n", i, s.superfluid_stiffness, s.dc_conductivity)); flush(f_trans[c])
            bn = bins[c]
            if bn[:count] == 0
                bn[:opt] = copy(s.optical_conductivity); bn[:dos] = copy(s.dos); bn[:dosAN] = copy(s.dos_AN)
                bn[:Ak0] = copy(s.A_k_ω0); bn[:count] = 1
            else
                bn[:opt] .+= s.optical_conductivity; bn[:dos] .+= s.dos; bn[:dosAN] .+= s.dos_AN; bn[:Ak0] .+= s.A_k_ω0
                bn[:count] += 1
            end
            if bn[:count] >= bin_size
                k = bn[:count]
                jldopen(jld[c], "a+") do file
                    g = JLD2.Group(file, "sweep_$i")
                    g["opt_cond"] = bn[:opt] ./ k; g["dos"] = bn[:dos] ./ k; g["dos_AN"] = bn[:dosAN] ./ k
                    g["A_k0"] = bn[:Ak0] ./ k; g["count"] = k
                end
                bn[:count] = 0
            end
        end
    end
    if i % 10 == 0
        for c in 1:B
            tee(c, @sprintf("Meas %d/%d. Acc=%.2f. E=%.4f", i, n_measure, acc_total[c] / i, obs[1, c]))
        end
    end
end
foreach(close, f_log); foreach(close, f_obs); foreach(close, f_trans)
