"""qinfo.py — host-side mirror of the QInfo readouts a quantum reservoir uses, and their batched device versions.

partial_trace / entanglement_entropy take the device state directly (the reference needs the full
density matrix rho = psi psi', 16 * 4^n bytes — QSim.jl:39-41, QInfo.jl:381-405; the engine
reduces psi straight to the 2^k x 2^k matrix in one read of the state).  The entropy functions run
on that small matrix on the host, exactly as QInfo.jl does (eigen-decomposition, 1e-10 clamps).
"""
import numpy as np

from .qsim import B200Density, B200State


def partial_trace(s, dim_A, dim_B, subsystem, batch_idx=0):
    """partial_trace(statevector_to_density_matrix(s), dim_A, dim_B, subsystem), QInfo.jl:381-405.
    A is the high-order factor (row index = a * dim_B + b).  subsystem == 1 traces out A and returns
    the dim_B x dim_B matrix; any other value traces out B (QInfo.jl:384, :392)."""
    assert isinstance(s, B200State)
    if dim_A * dim_B != (1 << s.n):
        raise AssertionError("Dimension mismatch")            # QInfo.jl:382
    kA, kB = int(round(np.log2(dim_A))), int(round(np.log2(dim_B)))
    if (1 << kA) != dim_A or (1 << kB) != dim_B:
        raise AssertionError("Dimension mismatch: dimensions must be powers of two on this engine")
    if subsystem == 1:
        return s.reduced_density(True, kB, batch_idx)
    return s.reduced_density(False, kA, batch_idx)


def matrix_log2(rho):
    """QInfo.jl:24-28"""
    w, v = np.linalg.eigh(rho)
    w = np.maximum(w, 1e-10)
    return v @ np.diag(np.log2(w)) @ v.conj().T


def von_neumann_entropy(rho):
    """QInfo.jl:47-50"""
    rho = (rho + rho.conj().T) / 2
    return float(-np.real(np.trace(rho @ matrix_log2(rho))))


def entanglement_entropy(s, dim_A, dim_B, batch_idx=0):
    """QInfo.jl:178-182 with rho_AB = psi psi'"""
    return von_neumann_entropy(partial_trace(s, dim_A, dim_B, 2, batch_idx))


def renyi_entropy(rho, alpha):
    """QInfo.jl:199-204"""
    assert alpha != 1.0, "Rényi entropy undefined for α=1; use von Neumann entropy"
    rho = (rho + rho.conj().T) / 2
    w, v = np.linalg.eigh(rho)
    rho_a = (v * np.power(w.astype(np.complex128), alpha)) @ v.conj().T
    return float((1 / (1 - alpha)) * np.log2(np.real(np.trace(rho_a))))


def max_entropy(rho):
    """QInfo.jl:332-337"""
    rho = (rho + rho.conj().T) / 2
    w = np.linalg.eigvalsh(rho)
    return float(np.log2(np.sum(w > 1e-10)))


def min_entropy(rho):
    """QInfo.jl:353-357"""
    rho = (rho + rho.conj().T) / 2
    return float(-np.log2(np.max(np.linalg.eigvalsh(rho))))


# ---- the remaining QInfo measures: host post-processing of SMALL matrices, as in the reference.
# They take numpy matrices — a reduced density matrix returned by the engine (partial_trace / reduced_density) or
# B200Density.to_numpy() — so that every QInfo function of the reference has its counterpart.

def _herm(rho):
    rho = np.asarray(rho, dtype=np.complex128)
    return (rho + rho.conj().T) / 2


def _herm_pow(rho, alpha):
    w, v = np.linalg.eigh(rho)
    return (v * np.power(w.astype(np.complex128), alpha)) @ v.conj().T


def partial_trace_matrix(rho, dim_A, dim_B, subsystem):
    """partial_trace(rho, dim_A, dim_B, subsystem) of a host matrix, QInfo.jl:381-405 (A = high-order factor)"""
    rho = np.asarray(rho, dtype=np.complex128)
    if rho.shape != (dim_A * dim_B, dim_A * dim_B):
        raise AssertionError("Dimension mismatch")
    r = _herm(rho).reshape(dim_A, dim_B, dim_A, dim_B)
    return np.einsum("kikj->ij", r) if subsystem == 1 else np.einsum("iljl->ij", r)


def _joint(x):
    """a joint state given as a host matrix, a B200Density, or a (pure) B200State"""
    if isinstance(x, B200Density):
        return x.to_numpy()
    if isinstance(x, B200State):
        v = x.to_numpy()
        return np.outer(v, v.conj())
    return np.asarray(x, dtype=np.complex128)


def quantum_mutual_information(rho_AB, dim_A, dim_B):
    """QInfo.jl:68-73: S(A) + S(B) - S(AB)"""
    r = _herm(_joint(rho_AB))
    return (von_neumann_entropy(partial_trace_matrix(r, dim_A, dim_B, 2)) + von_neumann_entropy(partial_trace_matrix(r, dim_A, dim_B, 1))
            - von_neumann_entropy(r))


def conditional_quantum_entropy(rho_AB, dim_A, dim_B):
    """QInfo.jl:91-95: S(AB) - S(B)"""
    r = _herm(_joint(rho_AB))
    return von_neumann_entropy(r) - von_neumann_entropy(partial_trace_matrix(r, dim_A, dim_B, 1))


def coherent_information(rho_AB, dim_A, dim_B):
    """QInfo.jl:113-117: S(B) - S(AB)"""
    r = _herm(_joint(rho_AB))
    return von_neumann_entropy(partial_trace_matrix(r, dim_A, dim_B, 1)) - von_neumann_entropy(r)


def relative_entropy(rho, sigma):
    """QInfo.jl:134-138: Tr(rho log2 rho) - Tr(rho log2 sigma)"""
    rho, sigma = _herm(rho), _herm(sigma)
    return float(np.real(np.trace(rho @ matrix_log2(rho)) - np.trace(rho @ matrix_log2(sigma))))


def holevo_information(states, probs):
    """QInfo.jl:155-160"""
    assert abs(sum(probs) - 1.0) < 1e-10, "Probabilities must sum to 1"
    avg = sum(p * np.asarray(r, dtype=np.complex128) for p, r in zip(probs, states))
    return von_neumann_entropy(avg) - sum(p * von_neumann_entropy(np.asarray(r, dtype=np.complex128)) for p, r in zip(probs, states))


def quantum_relative_renyi_entropy(rho, sigma, alpha):
    """QInfo.jl:222-227"""
    assert alpha != 1.0, "Relative Rényi entropy undefined for α=1"
    rho, sigma = _herm(rho), _herm(sigma)
    return float((1 / (alpha - 1)) * np.log2(np.real(np.trace(_herm_pow(rho, alpha) @ _herm_pow(sigma, 1 - alpha)))))


def quantum_fisher_information(psi, dpsi):
    """QInfo.jl:244-246: 4 (<dpsi|dpsi> - |<psi|dpsi>|^2)"""
    psi, dpsi = np.asarray(psi, dtype=np.complex128), np.asarray(dpsi, dtype=np.complex128)
    return float(4 * np.real(np.vdot(dpsi, dpsi) - abs(np.vdot(psi, dpsi)) ** 2))


def l1_norm_coherence(rho):
    """QInfo.jl:262-272: sum of |rho_ij| over i != j"""
    r = _herm(_joint(rho))
    return float(np.sum(np.abs(r)) - np.sum(np.abs(np.diag(r))))


def fidelity(rho, sigma):
    """QInfo.jl:289-294: (Tr sqrt(sqrt(rho) sigma sqrt(rho)))^2"""
    rho, sigma = _herm(_joint(rho)), _herm(_joint(sigma))
    sr = _herm_pow(rho, 0.5)
    w = np.linalg.eigvalsh(_herm(sr @ sigma @ sr))
    return float(np.real(np.sum(np.sqrt(w.astype(np.complex128)))) ** 2)


def trace_distance(rho, sigma):
    """QInfo.jl:311-316: half the sum of |eigenvalues| of rho - sigma"""
    d = _herm(_joint(rho)) - _herm(_joint(sigma))
    return float(0.5 * np.sum(np.abs(np.linalg.eigvalsh(d))))


# ---- batched measures on the device -------------------------------------------------------------
# One warp of k_herm_measure diagonalises one matrix (cyclic complex Jacobi); thousands of matrices are one launch.
# rhos / sigmas: arrays [count, d, d] (numpy [i, j]), d <= 16.

_KINDS = {"von_neumann": 0, "renyi": 1, "min": 2, "max": 3, "trace_distance": 4, "fidelity": 5, "eigvals": 6}


def _colmajor(ms):
    ms = np.asarray(ms, dtype=np.complex128)
    if ms.ndim == 2:
        ms = ms[None]
    assert ms.ndim == 3 and ms.shape[1] == ms.shape[2] and ms.shape[1] <= 16
    return np.ascontiguousarray(ms.transpose(0, 2, 1)), ms.shape[0], ms.shape[1]


def measure_batch(kind, rhos, sigmas=None, alpha=2.0, device=0):
    from . import _lib as L
    A, count, d = _colmajor(rhos)
    B = None
    if sigmas is not None:
        B, cb, db = _colmajor(sigmas)
        assert (cb, db) == (count, d)
    out = np.empty((count, d) if kind == "eigvals" else count, dtype=np.float64)
    L.check(L.lib().qm_herm_measure_batch(int(device), _KINDS[kind], float(alpha), L.dptr(A.view(np.float64)),
                                          L.dptr(B.view(np.float64)) if B is not None else None, d, count, L.dptr(out)))
    return out


def von_neumann_entropy_batch(rhos, device=0): return measure_batch("von_neumann", rhos, device=device)
def renyi_entropy_batch(rhos, alpha, device=0): return measure_batch("renyi", rhos, alpha=alpha, device=device)
def min_entropy_batch(rhos, device=0): return measure_batch("min", rhos, device=device)
def max_entropy_batch(rhos, device=0): return measure_batch("max", rhos, device=device)
def trace_distance_batch(rhos, sigmas, device=0): return measure_batch("trace_distance", rhos, sigmas, device=device)
def fidelity_batch(rhos, sigmas, device=0): return measure_batch("fidelity", rhos, sigmas, device=device)
def eigvalsh_batch(rhos, device=0): return measure_batch("eigvals", rhos, device=device)


def quantum_mutual_information_batch(rhos_AB, dim_A, dim_B, device=0):
    """QInfo.jl:68-73 for a batch of joint matrices [count, dA dB, dA dB] (dA dB <= 16): S(A) + S(B) - S(AB); the partial
    traces are reshapes on the host, the three eigen-decompositions per item run on the device"""
    r = np.asarray(rhos_AB, dtype=np.complex128)
    r = (r + r.conj().transpose(0, 2, 1)) / 2
    t = r.reshape(-1, dim_A, dim_B, dim_A, dim_B)
    rA, rB = np.einsum("niljl->nij", t), np.einsum("nkikj->nij", t)
    return (von_neumann_entropy_batch(rA, device) + von_neumann_entropy_batch(rB, device) - von_neumann_entropy_batch(r, device))


def entanglement_entropy_batch(s, dim_A, dim_B):
    """entanglement_entropy (QInfo.jl:178-182) of EVERY register of a batched B200State, on the device"""
    kA = int(round(np.log2(dim_A)))
    if dim_A * dim_B != (1 << s.n) or (1 << kA) != dim_A:
        raise AssertionError("Dimension mismatch")
    return s.entropy_batch(False, kA, "von_neumann")
