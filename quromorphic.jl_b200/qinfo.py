"""qinfo.py — host-side mirror of the QInfo readouts a quantum reservoir uses.

partial_trace / entanglement_entropy take the device state directly (the reference needs the full
density matrix rho = psi psi', 16 * 4^n bytes — QSim.jl:39-41, QInfo.jl:381-405; the engine
reduces psi straight to the 2^k x 2^k matrix in one read of the state).  The entropy functions run
on that small matrix on the host, exactly as QInfo.jl does (eigen-decomposition, 1e-10 clamps).
"""
import numpy as np

from .qsim import B200State


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
