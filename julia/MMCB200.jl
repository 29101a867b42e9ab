# MMCB200.jl — Julia binding of libmmc_b200.so (include/mmc_b200.h) for the reference driver
# (BradenDKelly/MetropolisMonteCarlo, Ewald/main.jl).  UNTESTED HERE: Julia is not installed in
# the build image; the same C ABI is exercised through Python ctypes (metropolismontecarlo_b200/
# _lib.py) by the parity tests.  Memory layouts are the reference's own (SURVEY.md A.6): soa.coords /
# moa.COM (Vector{SVector{3,Float64}}) go through with no copy.  Array arguments are declared Ptr{Cvoid} and
# the ARRAYS (not a) are handed to ccall: Base.cconvert then keeps them rooted for the duration of
# the call (no GC.@preserve needed) and accepts any element type, SVector included; a single SVector (a COM)
# travels as Ref(x), which unsafe_convert(Ptr{Cvoid}, ::RefValue) supports for isbits types.
module MMCB200

const LIB = get(ENV, "MMC_B200_LIB", "libmmc_b200")
const EWALD, WOLF, LJ_ONLY, LJ_ATOMS = Cint(0), Cint(1), Cint(2), Cint(3)

struct Config
    device::Cint; rank::Cint; world::Cint; sync_mode::Cint; stream::Ptr{Cvoid}
end

mutable struct Props            # mmc_properties  (Properties of Ewald/auxillary.jl:37-45 + components)
    energy::Cdouble; virial::Cdouble; coulomb::Cdouble
    lj::Cdouble; real::Cdouble; recip::Cdouble; self::Cdouble; wolf_const::Cdouble
    overlaps::Int64
    intra::Cdouble              # intramolecular correction, 0 unless set_intramolecular!(e, true)
    Props() = new(0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
end

mutable struct TrialResult      # mmc_trial_result
    lj_old::Cdouble; lj_vir_old::Cdouble; lj_new::Cdouble; lj_vir_new::Cdouble
    qq_old::Cdouble; qq_vir_old::Cdouble; qq_new::Cdouble; qq_vir_new::Cdouble
    d_recip::Cdouble; overlap_old::Cint; overlap_new::Cint
    TrialResult() = new(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
end

struct Engine
    h::Ptr{Cvoid}
end

function check(e::Engine, rc::Cint)
    rc < 0 && error("libmmc_b200: " * unsafe_string(ccall((:mmc_last_error, LIB), Cstring, (Ptr{Cvoid},), e.h)))
    rc
end

function Engine(; device = 0, rank = 0, world = 1, sync_mode = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    cfg = Ref(Config(device, rank, world, sync_mode, C_NULL))
    rc = ccall((:mmc_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h)
    rc != 0 && error("mmc_create: " * unsafe_string(ccall((:mmc_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    Engine(h[])
end
destroy(e::Engine) = ccall((:mmc_destroy, LIB), Cint, (Ptr{Cvoid},), e.h)

# soa, moa, vdwTable exactly as main.jl holds them (MakeAtomArrays "kmc", MakeTables)
function upload!(e::Engine, soa, moa, vdwTable, box, rc_lj, rc_qq)
    nt = size(vdwTable.ϵᵢⱼ, 1)
    check(e, ccall((:mmc_upload_system, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
         Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble),
        e.h, length(moa), length(soa), soa.coords, soa.charge, soa.atype,
        moa.firstAtom, moa.lastAtom, moa.COM, nt,
        vdwTable.ϵᵢⱼ, vdwTable.σᵢⱼ, box, rc_lj, rc_qq))
end

# PrepareEwaldVariables(ewald, box)  — Ewald/ewalds.jl:45-103
function prepare_ewald!(e::Engine, kappa, nk, k_sq_max, factor)
    n = Ref{Cint}(0)
    check(e, ccall((:mmc_ewald_prepare, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cint, Cint, Cdouble, Ref{Cint}),
                   e.h, kappa, nk, k_sq_max, factor, n))
    Int(n[])
end

# LJ_poly_ΔU(i, moa, soa, vdwTable, r_cut, box)  — Ewald/energy.jl:209-290
function LJ_poly_ΔU(e::Engine, i::Int)
    p = Ref{Cdouble}(0); v = Ref{Cdouble}(0)
    check(e, ccall((:mmc_lj_mol, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Cdouble}, Ref{Cdouble}), e.h, i, p, v))
    p[], v[]
end

# EwaldShort(i, moa, soa, sim_props, ewald, box)  — Ewald/ewalds.jl:892-910
function EwaldShort(e::Engine, i::Int)
    en = Ref{Cdouble}(0); v = Ref{Cdouble}(0); o = Ref{Cint}(0)
    check(e, ccall((:mmc_ewald_short, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Cdouble}, Ref{Cdouble}, Ref{Cint}), e.h, i, en, v, o))
    en[], v[], o[] != 0
end

# RecipMove(box, ewald, r_old, r_new, q)  — Ewald/ewalds.jl:718-826
function RecipMove(e::Engine, r_old::Vector, r_new::Vector, q::Vector{Float64})
    d = Ref{Cdouble}(0)
    check(e, ccall((:mmc_recip_move, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Cdouble}),
                   e.h, r_old, r_new, q, length(q), d))
    d[]
end
recip_commit!(e::Engine) = check(e, ccall((:mmc_recip_commit, LIB), Cint, (Ptr{Cvoid},), e.h))     # main.jl:621
recip_rollback!(e::Engine) = check(e, ccall((:mmc_recip_rollback, LIB), Cint, (Ptr{Cvoid},), e.h)) # main.jl:628

# in-place writes of main.jl:527,552 / 623-624
set_molecule!(e::Engine, i::Int, com, sites::Vector) =
    check(e, ccall((:mmc_set_molecule, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}), e.h, i, Ref(com), sites))

# page-lock long-lived host arrays (soa.coords, moa.COM) once, so uploads from them are true asynchronous DMA at full PCIe rate
host_register(a::Array) = ccall((:mmc_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), a, sizeof(a)) == 0 ||
    error("mmc_host_register failed")
host_unregister(a::Array) = ccall((:mmc_host_unregister, LIB), Cint, (Ptr{Cvoid},), a) == 0 ||
    error("mmc_host_unregister failed")

# all positions at once (bulk form of set_molecule!): soa.coords, moa.COM
upload_positions!(e::Engine, coords, com) =
    check(e, ccall((:mmc_upload_positions, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), e.h, coords, com))

# host arrays in, Properties out: upload_positions! + potential with the copies overlapped with compute
function potential_host(e::Engine, coords, com, style::Cint = EWALD)
    p = Props()
    check(e, ccall((:mmc_potential_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Props}),
                   e.h, coords, com, style, p))
    p
end

# bytes the last potential_host copied host -> device on this rank (a sharded engine copies only its slab's site blocks)
function last_host_bytes(e::Engine)
    n = Ref{Int64}(0)
    check(e, ccall((:mmc_last_host_bytes, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), e.h, n))
    n[]
end

# tuning / debugging switches (include/mmc_b200.h mmc_debug_set), e.g. debug_set(e, "rhok_kshard", 1)
debug_set(e::Engine, key::AbstractString, value::Integer) =
    check(e, ccall((:mmc_debug_set, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), e.h, key, value))

# opt-in intramolecular Ewald correction (the reference omits it, Ewald/energy.jl:1008-1021); Props.intra carries it
set_intramolecular!(e::Engine, on::Bool) = check(e, ccall((:mmc_set_intramolecular, LIB), Cint, (Ptr{Cvoid}, Cint), e.h, on ? 1 : 0))

# LJ_poly_ΔU(i, …) and EwaldShort(i, …)[1] for every molecule i from one evaluation (the rows potential() sums)
function energy_all(e::Engine, n_mol::Integer, style::Cint = EWALD)
    lj = Vector{Float64}(undef, n_mol); vir = similar(lj); qq = similar(lj); ov = Vector{Int32}(undef, n_mol)
    check(e, ccall((:mmc_energy_all, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                   e.h, style, lj, vir, qq, ov))
    lj, vir, qq, ov
end

# potential(moa, soa, tot, ewald, vdwTable, sim_props[, "ewald"])  — Ewald/energy.jl:864-1032
function potential(e::Engine, style::Cint = EWALD)
    p = Props()
    check(e, ccall((:mmc_potential, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Props}), e.h, style, p))
    p
end

# fused fast path: the five calls of Loop (main.jl:491-590) in one launch
function trial_move(e::Engine, i::Int, com_new, sites_new::Vector, style::Cint = EWALD)
    r = TrialResult()
    check(e, ccall((:mmc_trial_move, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{TrialResult}),
                   e.h, i, Ref(com_new), sites_new, style, r))
    r
end
accept!(e::Engine) = check(e, ccall((:mmc_accept, LIB), Cint, (Ptr{Cvoid},), e.h))
reject!(e::Engine) = check(e, ccall((:mmc_reject, LIB), Cint, (Ptr{Cvoid},), e.h))

# Ewald/volumeChange.jl:50-147
function volume_trial(e::Engine, box_new, kappa_new, style::Cint = EWALD)
    p = Props()
    check(e, ccall((:mmc_volume_trial, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cint, Ref{Props}), e.h, box_new, kappa_new, style, p))
    p
end
volume_accept!(e::Engine) = check(e, ccall((:mmc_volume_accept, LIB), Cint, (Ptr{Cvoid},), e.h))
volume_reject!(e::Engine) = check(e, ccall((:mmc_volume_reject, LIB), Cint, (Ptr{Cvoid},), e.h))

# sharded full energy with the exchange over NVLink peer memory (one process per GPU; include/mmc_b200.h mmc_peer_*)
function peer_export(e::Engine)
    hd = Vector{UInt8}(undef, 64)
    check(e, ccall((:mmc_peer_export, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.h, hd))
    hd
end
peer_import!(e::Engine, rank::Integer, hd::Vector{UInt8}) =
    check(e, ccall((:mmc_peer_import, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), e.h, rank, hd))
function potential_sharded(e::Engine, style::Cint = EWALD)
    p = Props()
    check(e, ccall((:mmc_potential_sharded, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Props}), e.h, style, p))
    p
end

# Loop() for a whole block of moves in one launch (Ewald/main.jl:487-651); include/mmc_b200.h mmc_loop_run_device.
# com::Vector{SVector{3,Float64}} (moa.COM), quat::Vector{SVector{4,Float64}}, db: body-fixed site vectors (n_sites x 3),
# u: the stretch of the caller's uniform stream this block may consume (st.uniforms_used tells how much it did).
struct LoopParams
    temperature::Cdouble; dr_max::Cdouble; dphi_max::Cdouble
    p_trans::Cdouble; p_rot::Cdouble
    style::Cint; adjust::Cint
end
mutable struct LoopStats
    n_moves::Int64; n_accepted::Int64; n_overlap::Int64; uniforms_used::Int64
    trans_attempt::Int64; trans_accept::Int64; rot_attempt::Int64; rot_accept::Int64
    dr_max::Cdouble; dphi_max::Cdouble; total_energy::Cdouble; total_virial::Cdouble
    LoopStats() = new(0, 0, 0, 0, 0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0)
end
function loop_run_device!(e::Engine, p::LoopParams, com, quat, db, u::Vector{Float64}, n_moves::Int, e0, v0;
                          accepted::Vector{UInt8} = Vector{UInt8}(undef, n_moves), delta::Vector{Float64} = Vector{Float64}(undef, n_moves))
    st = LoopStats()
    rc = ccall((:mmc_loop_run_device, LIB), Cint,
               (Ptr{Cvoid}, Ref{LoopParams}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble,
                Ptr{Cvoid}, Ptr{Cvoid}, Ref{LoopStats}),
               e.h, p, com, quat, db, u, length(u), n_moves, e0, v0, accepted, delta, st)
    rc < 0 && check(e, rc)          # 1 = stream ran dry, 2 = quaternion-norm error (quaternions.jl:22-25), 3 = no move selected
    st, accepted, delta, rc
end

end # module
