# B200QSim.jl — Julia host shim: the third "method family" of Quromorphic.QSim / QInfo.
#
# The reference dispatches every function on the state's concrete type and ships two families,
# Vector{ComplexF64} and Matrix{ComplexF64} (src/Quantum/QSim.jl:52-383).  This file adds a third,
# `B200State`, whose methods are one `ccall` each into libqm_b200.so (include/qm_b200.h).  User code
# written against QSim keeps working after `s = statevector(n, m)` is replaced by
# `s = B200State(n, m)`.
#
# STATUS: Julia is not installed in the build image or on the GPU box, so this file has NEVER RUN UNDER A JULIA RUNTIME.
# It is parsed and its host logic executed by the Julia-subset interpreter oracle/jl_subset.py on top of the reference's own
# source, with ccall recorded: every ccall is checked against include/qm_b200.h (tests/test_julia_shim_static.py) and the calls
# it makes are compared one by one with the Python mirror's (tests/test_julia_shim_exec.py).  The same ABI is exercised by the Python
# mirror quromorphic.jl_b200/qsim.py, which is what tests/ run.  Every conversion below (1-based
# qubit -> 0-based bit, the reference's control map, argument checks and their order) is the same
# code as in qsim.py, where it is tested.
module B200QSim

using Quromorphic.QSim: H, X, Y, Z, I2, rx_gate, ry_gate, rz_gate
import Quromorphic.QSim: nb, u2!, h!, x!, y!, z!, rx!, ry!, rz!, controlled!, cnot!, crx!, cry!, crz!,
                          swap!, mp, measure_z, measure_x, measure_y, prstate,
                          statevector_to_density_matrix                      # new METHODS of the reference's functions
# (not `density_matrix`: its reference method takes (::Int, ::Int) and nothing of ours — a method with that signature would
#  REPLACE the reference's own, keyword arguments do not take part in dispatch; `B200Density(q, m)` is the constructor instead,
#  as `B200State(n, m)` stands for `statevector(n, m)`.  Found by executing this file: tests/test_julia_shim_exec.py)
import Quromorphic.QInfo: partial_trace, entanglement_entropy, von_neumann_entropy

export B200State, B200Density, reset!, sync!, norm2, apply_batched!, expect_z_all_begin!, expect_z_all_end!, reduced_density_batch, entropy_batch, dist_unique_id, dist_ipc_export, dist_ipc_import!, expect_z_all, expect_all, reduced_density, sample, apply_circuit!, apply_circuit_expect_z!, precompile_circuit!, jit_cache_dir, QMGate

const LIB = get(ENV, "QM_B200_LIB", joinpath(@__DIR__, "..", "libqm_b200.so"))

# struct qm_gate (80 bytes): kind, q0, q1, flags, m[8]
struct QMGate
    kind::Int32
    q0::Int32
    q1::Int32
    flags::Int32
    m::NTuple{8,Float64}
end

mutable struct B200State
    handle::Ptr{Cvoid}
    n::Int
    batch::Int
    mode::Symbol          # :reference reproduces controlled!'s index map (QSim.jl:181-197); :corrected does not
    threshold::Float64    # measure_z's probability cut (QSim.jl:308)
    world::Int            # ranks the register is sharded over (1: the whole register is on this GPU)
end

function lasterror()
    buf = Vector{UInt8}(undef, 512)
    ccall((:qm_last_error, LIB), Int32, (Ptr{UInt8}, Int32), buf, 512)
    unsafe_string(pointer(buf))
end

check(rc) = rc == 0 ? nothing : error(lasterror())      # ErrorException, like the reference

# qm_set_option keys (include/qm_b200.h).  The defaults are the measured ones; these are tuning / debugging knobs.
const OPT_EAGER, OPT_TILE_BITS, OPT_LOW_BITS, OPT_FUSE, OPT_TRACE, OPT_PIPELINE, OPT_MAX_COST, OPT_TIME_PASSES, OPT_SHORT_PCT = 0:8
const OPT_OVERLAP, OPT_XSM, OPT_XUN, OPT_SLAB_BITS, OPT_OVERLAP_DEPTH, OPT_JIT = 9:14      # OPT_JIT: compiled pass kernels (0 never, 1 at first sight, default 2)
const OPT_DEFER_PASS, OPT_ZFOLD = 15, 16                                                    # OPT_ZFOLD: <Z> sums inside the last pass of apply_circuit_expect_z! (default 1)
set_option!(s, key::Integer, value::Integer) =
    check(ccall((:qm_set_option, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), s.handle, key, value))

"""
    B200State(n, m = 0; batch = 1, device = 0, mode = :reference, threshold = 1e-10)

`statevector(n, m)` (QSim.jl:33-37) on the GPU.
"""
function B200State(n::Int, m::Int = 0; batch::Int = 1, device::Int = 0, mode::Symbol = :reference,
                   threshold::Float64 = 1e-10)
    0 <= m < (1 << n) || throw(BoundsError())            # statevector(n, m) indexes s[m+1]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qm_create, LIB), Int32, (Int32, UInt64, Int32, Int32, Ptr{Ptr{Cvoid}}), n, m, batch, device, h))
    st = B200State(h[], n, batch, mode, threshold, 1)
    finalizer(s -> (s.handle != C_NULL && ccall((:qm_destroy, LIB), Int32, (Ptr{Cvoid},), s.handle); s.handle = C_NULL), st)
    st
end

# ---- sharded registers: one Julia process per GPU (include/qm_b200.h, "sharded registers") ----------------------------------
"128-byte id of a sharded register: rank 0 calls this and hands it to the other ranks (MPI.bcast, Distributed, a file ...)"
function dist_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:qm_dist_unique_id, LIB), Int32, (Ptr{UInt8},), id))
    id
end

"""
    B200State(n, rank, world, id; m = 0, device = rank, mode = :reference, threshold = 1e-10)

This rank's shard of an `n`-qubit register sharded over `world` GPUs (a power of two): the top log2(world) qubits select the
rank.  Every gate and readout call of this file then works on the GLOBAL register as a collective call — all ranks make the same
calls in the same order and get the same answers; `mp` and `Vector{ComplexF64}(st)` return this rank's contiguous slice.
"""
function B200State(n::Int, rank::Int, world::Int, id::Vector{UInt8}; m::Int = 0, device::Int = rank, mode::Symbol = :reference,
                   threshold::Float64 = 1e-10)
    length(id) == 128 || error("the id of a sharded register has 128 bytes (dist_unique_id)")
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qm_create_dist, LIB), Int32, (Int32, UInt64, Int32, Int32, Int32, Ptr{UInt8}, Ptr{Ptr{Cvoid}}),
                n, m, device, rank, world, id, h))
    st = B200State(h[], n, 1, mode, threshold, world)
    finalizer(s -> (s.handle != C_NULL && ccall((:qm_destroy, LIB), Int32, (Ptr{Cvoid},), s.handle); s.handle = C_NULL), st)
    st
end

"64 bytes every rank exports after creating its shard; all-gathered (world x 64 bytes, rank-major) they go to `dist_ipc_import!`"
function dist_ipc_export(s::B200State)
    hnd = Vector{UInt8}(undef, 64)
    check(ccall((:qm_dist_ipc_export, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), s.handle, hnd))
    hnd
end
"map the other ranks' shards into this process (peer memory over NVLink: the exchange runs beside the fused passes); false = the engine keeps NCCL"
dist_ipc_import!(s::B200State, handles::Vector{UInt8}) =
    ccall((:qm_dist_ipc_import, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), s.handle, handles) == 0

"upload: `B200State(v)`"
function B200State(v::Vector{ComplexF64}; kw...)
    n = round(Int, log2(length(v)))
    (1 << n) == length(v) || error("state length must be a power of two")
    st = B200State(n, 0; kw...)
    check(ccall((:qm_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, UInt64, Int32), st.handle, v, length(v), 0))
    st
end

"download: `Vector{ComplexF64}(st)` (a sharded register: this rank's contiguous slice)"
function Base.Vector{ComplexF64}(st::B200State; batch_idx::Int = 0)
    v = Vector{ComplexF64}(undef, (1 << st.n) ÷ st.world)
    check(ccall((:qm_download, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, UInt64, Int32), st.handle, v, length(v), batch_idx))
    v
end
Base.length(st::B200State) = 1 << st.n
nb(st::B200State) = st.n                                  # QSim.jl:48

# ---- gates ------------------------------------------------------------------------------------
# A Matrix{ComplexF64} 2x2 is already the 8 doubles the ABI wants (column-major, interleaved).
function u2!(s::B200State, t::Int, U::Matrix{ComplexF64})  # QSim.jl:52-63
    1 <= t <= s.n || error("Qubit index out of range")
    check(ccall((:qm_apply_u2, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{ComplexF64}), s.handle, t - 1, U))
    s
end
h!(s::B200State, t::Int) = u2!(s, t, H)
x!(s::B200State, t::Int) = u2!(s, t, X)
y!(s::B200State, t::Int) = u2!(s, t, Y)
z!(s::B200State, t::Int) = u2!(s, t, Z)
rx!(s::B200State, t::Int, theta::Real) = u2!(s, t, rx_gate(theta))
ry!(s::B200State, t::Int, theta::Real) = u2!(s, t, ry_gate(theta))
rz!(s::B200State, t::Int, theta::Real) = u2!(s, t, rz_gate(theta))

"what controlled!(s, c, t, V) really acts on (QSim.jl:181-197): qubits (b-1, b), b = max(c, t)"
effective_ct(c::Int, t::Int) = (b = max(c, t); c < t ? (b - 1, b) : (b, b - 1))

function controlled!(s::B200State, c::Int, t::Int, V::Matrix{ComplexF64})   # QSim.jl:175-203
    (c <= s.n && t <= s.n && c >= 1 && t >= 1) || error("Qubit index out of range")   # :177 (range first)
    c == t && error("Control and target cannot be the same qubit")                    # :178
    ce, te = s.mode === :reference ? effective_ct(c, t) : (c, t)
    check(ccall((:qm_apply_controlled, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{ComplexF64}), s.handle, ce - 1, te - 1, V))
    s
end
cnot!(s::B200State, c::Int, t::Int) = controlled!(s, c, t, X)
crx!(s::B200State, c::Int, t::Int, theta::Real) = controlled!(s, c, t, rx_gate(theta))
cry!(s::B200State, c::Int, t::Int, theta::Real) = controlled!(s, c, t, ry_gate(theta))
crz!(s::B200State, c::Int, t::Int, theta::Real) = controlled!(s, c, t, rz_gate(theta))

function swap!(s::B200State, q1::Int, q2::Int)            # QSim.jl:262-275
    (q1 <= s.n && q2 <= s.n && q1 >= 1 && q2 >= 1) || error("Qubit index out of range")
    q1 == q2 && return s
    cnot!(s, q1, q2); cnot!(s, q2, q1); cnot!(s, q1, q2)
    s
end

"""
    precompile_circuit!(s, gates) -> Int

Compile the pass kernels `gates` will need (QM_OPT_JIT) ahead of its first application; the register is not touched.
Only the structure of `gates` matters (gate kinds and qubits), not the angles.  `jit_cache_dir(path)` keeps the kernels
as files for later processes.
"""
function precompile_circuit!(s::B200State, gates::Vector{QMGate})
    k = Ref{Int32}(0)
    check(ccall((:qm_precompile_circuit, LIB), Int32, (Ptr{Cvoid}, Ptr{QMGate}, Int32, Ptr{Int32}), s.handle, gates, length(gates), k))
    return Int(k[])
end
jit_cache_dir(path::AbstractString) = check(ccall((:qm_jit_cache_dir, LIB), Int32, (Cstring,), path))

"""
    apply_circuit_expect_z!(s, gates; threshold = s.threshold) -> Array{Float64} (2 x n x batch: out[:, q, b] = measure_z of qubit q of register b)

A reservoir step in one call: `apply_circuit!` followed by `expect_z_all`.  On a large register the sums of `measure_z` are
taken inside the last fused pass (QM_OPT_ZFOLD), so the state is not read again for the readout.
"""
function apply_circuit_expect_z!(s::B200State, gates::Vector{QMGate}; threshold::Float64 = s.threshold)
    out = Array{Float64}(undef, 2, s.n, s.batch)
    check(ccall((:qm_apply_circuit_expect_z, LIB), Int32, (Ptr{Cvoid}, Ptr{QMGate}, Int32, Float64, Ptr{Float64}),
                s.handle, gates, length(gates), threshold, out))
    return out
end

"apply a whole gate list (0-based textbook bits, the wire format) in as few HBM passes as possible"
function apply_circuit!(s::B200State, gates::Vector{QMGate})
    check(ccall((:qm_apply_circuit, LIB), Int32, (Ptr{Cvoid}, Ptr{QMGate}, Int32), s.handle, gates, length(gates)))
    s
end

# ---- what a reservoir loop needs beyond the reference's own API (no reference counterpart; same calls as the Python mirror) ----
"back to |m> for every register of the batch (the reference allocates a new statevector(n, m) instead)"
reset!(s::B200State, m::Int = 0) = (check(ccall((:qm_reset, LIB), Int32, (Ptr{Cvoid}, UInt64), s.handle, m)); s)
"wait for everything queued on the register"
sync!(s::B200State) = (check(ccall((:qm_sync, LIB), Int32, (Ptr{Cvoid},), s.handle)); s)
"sum of abs2 of register `batch_idx` (no threshold)"
function norm2(s::B200State; batch_idx::Int = 0)
    v = Ref{Float64}(0.0)
    check(ccall((:qm_norm2, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), s.handle, batch_idx, v))
    v[]
end

"""
    apply_batched!(s, gates, mats)

One gate list for every register of the batch; a gate of kind QM_GATE_U2_SLOT takes ITS register's matrix from
`mats[:, slot, b]` (8 x nslots x batch Float64: the 8 doubles of a Matrix{ComplexF64} 2x2, e.g. a per-register ry encoding).
"""
function apply_batched!(s::B200State, gates::Vector{QMGate}, mats::Array{Float64,3})
    size(mats, 1) == 8 && size(mats, 3) == s.batch || error("mats must be 8 x nslots x batch")
    check(ccall((:qm_apply_batched, LIB), Int32, (Ptr{Cvoid}, Ptr{QMGate}, Int32, Ptr{Float64}, Int32),
                s.handle, gates, length(gates), mats, size(mats, 2)))
    s
end

"first half of a split readout: enqueue the all-qubit measure_z of the state as it is NOW and return at once (at most two pending)"
expect_z_all_begin!(s::B200State; threshold::Float64 = s.threshold) =
    (check(ccall((:qm_expect_z_all_begin, LIB), Int32, (Ptr{Cvoid}, Float64), s.handle, threshold)); s)
"second half: the oldest pending readout, 2 x n x batch"
function expect_z_all_end!(s::B200State)
    out = Array{Float64}(undef, 2, s.n, s.batch)
    check(ccall((:qm_expect_z_all_end, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), s.handle, out))
    out
end

"reduced density matrix of EVERY register of the batch: 2^k x 2^k x batch (QInfo.partial_trace of s*s', per register)"
function reduced_density_batch(s::B200State, keep_low::Bool, k::Int)
    d = 1 << k
    out = Array{ComplexF64}(undef, d, d, s.batch)
    check(ccall((:qm_reduced_density_batch, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{ComplexF64}), s.handle, keep_low ? 1 : 0, k, out))
    out
end
"entropy of the reduced matrix of every register, on the device: kind = :von_neumann, :renyi (alpha), :min, :max (QInfo.jl:47-50, :199-204, :353-357, :332-337)"
function entropy_batch(s::B200State, keep_low::Bool, k::Int; kind::Symbol = :von_neumann, alpha::Float64 = 2.0)
    code = kind === :von_neumann ? 0 : kind === :renyi ? 1 : kind === :min ? 2 : kind === :max ? 3 : error("unknown entropy kind")
    out = Vector{Float64}(undef, s.batch)
    check(ccall((:qm_entropy_batch, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Float64, Ptr{Float64}), s.handle, keep_low ? 1 : 0, k, code, alpha, out))
    out
end

"read a gate-list file written by circuits.save_gates (binary form: b\"QMGATES1\" | u32 n | u32 ngates | 80-byte records)"
function load_gates(path::AbstractString)
    open(path, "r") do io
        read(io, 8) == Vector{UInt8}("QMGATES1") || error("not a QMGATES1 file")
        n = Int(ltoh(read(io, UInt32))); ng = Int(ltoh(read(io, UInt32)))
        gates = Vector{QMGate}(undef, ng)
        read!(io, gates)
        (n, gates)
    end
end

# ---- readouts -----------------------------------------------------------------------------------
function mp(s::B200State; batch_idx::Int = 0)             # QSim.jl:293 (a sharded register: this rank's slice)
    p = Vector{Float64}(undef, (1 << s.n) ÷ s.world)
    check(ccall((:qm_probs, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32), s.handle, p, batch_idx))
    p
end

function _measure(s::B200State, t::Int, R::Union{Nothing,Matrix{ComplexF64}}, batch_idx::Int)
    1 <= t <= s.n || error("Qubit index out of range")
    out = Vector{Float64}(undef, 2)
    Rp = R === nothing ? Ptr{ComplexF64}(C_NULL) : pointer(R)
    GC.@preserve R check(ccall((:qm_measure_with, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{ComplexF64}, Float64, Int32, Ptr{Float64}),
                               s.handle, t - 1, Rp, s.threshold, batch_idx, out))
    (out[1], out[2])
end
measure_z(s::B200State, t::Int; batch_idx::Int = 0) = _measure(s, t, nothing, batch_idx)          # QSim.jl:297-318
measure_x(s::B200State, t::Int; batch_idx::Int = 0) = _measure(s, t, H, batch_idx)                # QSim.jl:320-324
measure_y(s::B200State, t::Int; batch_idx::Int = 0) = _measure(s, t, rx_gate(π / 2), batch_idx)   # QSim.jl:326-330

"all-qubit measure_z in one pass: 2 x n x batch array of (p0, p1)"
function expect_z_all(s::B200State; threshold::Float64 = s.threshold)
    out = Array{Float64}(undef, 2, s.n, s.batch)
    check(ccall((:qm_expect_z_all, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}), s.handle, threshold, out))
    out
end

"""
    expect_all(s, basis)   basis = :z, :x, :y or a 2x2 rotation matrix

measure_z / measure_x / measure_y (QSim.jl:297-330) of EVERY qubit: 2 x n x batch array of (p0, p1).  X and Y read the
state once per window of 12 (then 9) qubits and rotate every pair in registers; the matrices are the host's own H and
rx_gate(pi/2), so the engine sees the same bits measure_x / measure_y pass.
"""
function expect_all(s::B200State, basis; threshold::Float64 = s.threshold)
    R = basis === :z ? nothing : basis === :x ? H : basis === :y ? rx_gate(π / 2) : basis
    out = Array{Float64}(undef, 2, s.n, s.batch)
    Rp = R === nothing ? Ptr{ComplexF64}(C_NULL) : pointer(R)
    GC.@preserve R check(ccall((:qm_expect_all_with, LIB), Int32, (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Ptr{Float64}),
                               s.handle, Rp, threshold, out))
    out
end

function prstate(s::B200State)                            # QSim.jl:367-373, after a download
    v = Vector{ComplexF64}(s)
    for i in 0:(length(v) - 1)
        a = v[i + 1]
        abs(a) > 1e-10 && println("|", lpad(string(i, base = 2), s.n, '0'), "⟩: ", a)
    end
end

# ---- QInfo --------------------------------------------------------------------------------------
"2^k x 2^k reduced density matrix straight from psi; keep_low keeps B (the low k qubits)"
function reduced_density(s::B200State, keep_low::Bool, k::Int; batch_idx::Int = 0)
    d = 1 << k
    rho = Matrix{ComplexF64}(undef, d, d)                 # column-major, as the ABI writes it
    check(ccall((:qm_reduced_density, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{ComplexF64}), s.handle, keep_low ? 1 : 0, k, batch_idx, rho))
    rho
end

"partial_trace(statevector_to_density_matrix(s), dim_A, dim_B, subsystem) without forming rho (QInfo.jl:381-405)"
function partial_trace(s::B200State, dim_A::Int, dim_B::Int, subsystem::Int)
    @assert dim_A * dim_B == (1 << s.n) "Dimension mismatch"
    kA, kB = round(Int, log2(dim_A)), round(Int, log2(dim_B))
    @assert (1 << kA) == dim_A && (1 << kB) == dim_B "Dimension mismatch"
    subsystem == 1 ? reduced_density(s, true, kB) : reduced_density(s, false, kA)
end
entanglement_entropy(s::B200State, dim_A::Int, dim_B::Int) = von_neumann_entropy(partial_trace(s, dim_A, dim_B, 2))

function sample(s::B200State, seed::Integer, nshots::Int; batch_idx::Int = 0)
    out = Vector{UInt64}(undef, nshots)
    check(ccall((:qm_sample, LIB), Int32, (Ptr{Cvoid}, UInt64, Int32, Int32, Ptr{UInt64}), s.handle, seed, nshots, batch_idx, out))
    out
end

# ---- density matrices: the Matrix{ComplexF64} methods (QSim.jl:67-79, :209-232, :277-290, :295, :332-365) ----
# rho on q qubits is a 2q-qubit register holding vec(rho) (column-major, as Julia stores the Matrix):
# rho <- op rho op' is `op` on the row bits and conj(op) on the column bits.
struct B200Density
    vec::B200State          # the 2q-qubit register (mode = :corrected: the index map is applied here)
    q::Int
    mode::Symbol
end
B200Density(q::Int, m::Int = 0; mode::Symbol = :reference, kw...) =
    B200Density(B200State(2q, m * ((1 << q) + 1); mode = :corrected, kw...), q, mode)      # density_matrix(q, m), QSim.jl:43-46
function statevector_to_density_matrix(s::B200State)                                          # QSim.jl:39-41
    rho = B200Density(s.n, 0; mode = s.mode)
    check(ccall((:qm_dm_from_state, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), s.handle, rho.vec.handle))
    rho
end
nb(rho::B200Density) = rho.q
function u2!(rho::B200Density, t::Int, U::Matrix{ComplexF64})                                 # QSim.jl:67-79
    1 <= t <= rho.q || error("Qubit index out of range")
    Uc = conj(U)
    check(ccall((:qm_apply_u2, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{ComplexF64}), rho.vec.handle, t - 1, U))
    check(ccall((:qm_apply_u2, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{ComplexF64}), rho.vec.handle, rho.q + t - 1, Uc))
    rho
end
function controlled!(rho::B200Density, c::Int, t::Int, V::Matrix{ComplexF64})                 # QSim.jl:209-232
    (c <= rho.q && t <= rho.q && c >= 1 && t >= 1) || error("Qubit index out of range")
    c == t && error("Control and target cannot be the same qubit")
    ce, te = rho.mode === :reference ? effective_ct(c, t) : (c, t)
    Vc = conj(V)
    check(ccall((:qm_apply_controlled, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{ComplexF64}), rho.vec.handle, ce - 1, te - 1, V))
    check(ccall((:qm_apply_controlled, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{ComplexF64}), rho.vec.handle, rho.q + ce - 1, rho.q + te - 1, Vc))
    rho
end
for (f, U) in ((:h!, :H), (:x!, :X), (:y!, :Y), (:z!, :Z))
    @eval $f(rho::B200Density, t::Int) = u2!(rho, t, $U)
end
rx!(rho::B200Density, t::Int, theta::Real) = u2!(rho, t, rx_gate(theta))
ry!(rho::B200Density, t::Int, theta::Real) = u2!(rho, t, ry_gate(theta))
rz!(rho::B200Density, t::Int, theta::Real) = u2!(rho, t, rz_gate(theta))
cnot!(rho::B200Density, c::Int, t::Int) = controlled!(rho, c, t, X)
crx!(rho::B200Density, c::Int, t::Int, theta::Real) = controlled!(rho, c, t, rx_gate(theta))
cry!(rho::B200Density, c::Int, t::Int, theta::Real) = controlled!(rho, c, t, ry_gate(theta))
crz!(rho::B200Density, c::Int, t::Int, theta::Real) = controlled!(rho, c, t, rz_gate(theta))
function swap!(rho::B200Density, q1::Int, q2::Int)                                            # QSim.jl:277-290
    (q1 <= rho.q && q2 <= rho.q && q1 >= 1 && q2 >= 1) || error("Qubit index out of range")
    q1 == q2 && return rho
    cnot!(rho, q1, q2); cnot!(rho, q2, q1); cnot!(rho, q1, q2)
    rho
end
function mp(rho::B200Density)                                                                 # QSim.jl:295
    p = Vector{Float64}(undef, 1 << rho.q)
    check(ccall((:qm_dm_diag, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), rho.vec.handle, p))
    p
end
function _measure(rho::B200Density, t::Int, R::Union{Nothing,Matrix{ComplexF64}})
    1 <= t <= rho.q || error("Qubit index out of range")
    out = Vector{Float64}(undef, 2)
    Rp = R === nothing ? Ptr{ComplexF64}(C_NULL) : pointer(R)
    GC.@preserve R check(ccall((:qm_dm_measure, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{ComplexF64}, Float64, Ptr{Float64}),
                               rho.vec.handle, t - 1, Rp, rho.vec.threshold, out))
    (out[1], out[2])
end
measure_z(rho::B200Density, t::Int) = _measure(rho, t, nothing)          # QSim.jl:332-352
measure_x(rho::B200Density, t::Int) = _measure(rho, t, H)                # QSim.jl:354-358
measure_y(rho::B200Density, t::Int) = _measure(rho, t, rx_gate(π / 2))   # QSim.jl:360-364
Base.Matrix{ComplexF64}(rho::B200Density) = reshape(Vector{ComplexF64}(rho.vec), 1 << rho.q, 1 << rho.q)
function prstate(rho::B200Density)                                       # QSim.jl:375-383
    p = mp(rho)
    for i in 0:(length(p) - 1)
        abs(p[i + 1]) > 1e-10 && println("|", lpad(string(i, base = 2), rho.q, '0'), "⟩ : ", p[i + 1])
    end
    nothing
end

end # module
