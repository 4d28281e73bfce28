"""literal.py — TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy transcription of the reference's *literal* algorithm: build the full 2^n x 2^n operator
with kron and multiply it into the state, exactly as /root/reference/src/Quantum/QSim.jl does.
It exists to pin the matrix-free C oracle (oracle/qsim_oracle.c): Julia is not installed here, so
this is the closest executable statement of what the reference computes.  O(4^n) — n <= 10.

Qubit arguments are the reference's 1-based indices.  numpy's kron(A, B) has the same block
structure as Julia's kron(A, B) (A is the slow / high-order factor).
"""
import numpy as np

c1 = np.complex128(1)
im = np.complex128(1j)
# QSim.jl:10-17
I2 = np.array([[1, 0], [0, 1]], dtype=np.complex128)
H = np.complex128(1 / np.sqrt(2)) * np.array([[1, 1], [1, -1]], dtype=np.complex128)
X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -im], [im, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
P0 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
P1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)


def id_(n):
    """QSim.jl:29-31 (dense here; the reference returns Diagonal for n > 0)."""
    return np.eye(1 << n, dtype=np.complex128)


def statevector(n, m):
    """QSim.jl:33-37"""
    s = np.zeros(1 << n, dtype=np.complex128)
    s[m] = c1
    return s


def nb(s):
    """QSim.jl:48"""
    return int(round(np.log2(len(s))))


def rx_gate(theta):
    """QSim.jl:114-124 (without the memo cache): c*I2 - im*s*X"""
    c = np.complex128(np.cos(np.float64(theta) / 2))
    s = np.complex128(np.sin(np.float64(theta) / 2))
    return c * I2 - im * s * X


def ry_gate(theta):
    """QSim.jl:126-136"""
    c = np.complex128(np.cos(np.float64(theta) / 2))
    s = np.complex128(np.sin(np.float64(theta) / 2))
    return c * I2 - im * s * Y


def rz_gate(theta):
    """QSim.jl:138-148"""
    c = np.complex128(np.cos(np.float64(theta) / 2))
    s = np.complex128(np.sin(np.float64(theta) / 2))
    return c * I2 - im * s * Z


def u2(s, t, U):
    """QSim.jl:52-63: op = kron(id(r), kron(U, id(l))); s <- op * s"""
    q = nb(s)
    if not t <= q:
        raise RuntimeError("Qubit index out of range")
    l, r = t - 1, q - t
    op = np.kron(id_(r), np.kron(U, id_(l)))
    s[:] = op @ s
    return s


def controlled(s, c, t, V):
    """QSim.jl:175-203, transcribed operator for operator (this is what carries the
    non-adjacent behaviour: U2 sits directly above `mid`)."""
    q = nb(s)
    if not (c <= q and t <= q):
        raise RuntimeError("Qubit index out of range")
    if c == t:
        raise RuntimeError("Control and target cannot be the same qubit")
    a, b = min(c, t), max(c, t)
    left = id_(q - b)
    right = id_(a - 1)
    mid = id_(b - a - 1) if (b - a - 1) > 0 else np.array([[c1]])
    if c < t:
        U2 = np.kron(id_(1), P0) + np.kron(V, P1)
    else:
        U2 = np.kron(P0, id_(1)) + np.kron(P1, V)
    op = np.kron(left, np.kron(U2, np.kron(mid, right)))
    s[:] = op @ s
    return s


def cnot(s, c, t):
    return controlled(s, c, t, X)


def swap(s, q1, q2):
    """QSim.jl:262-275"""
    q = nb(s)
    if not (q1 <= q and q2 <= q):
        raise RuntimeError("Qubit index out of range")
    if q1 == q2:
        return s
    cnot(s, q1, q2)
    cnot(s, q2, q1)
    cnot(s, q1, q2)
    return s


def mp(s):
    """QSim.jl:293"""
    return s.real * s.real + s.imag * s.imag


def measure_z(s, t, threshold=1e-10):
    """QSim.jl:297-318 (sequential, thresholded)."""
    n = nb(s)
    if not t <= n:
        raise RuntimeError("Qubit index out of range")
    probs = mp(s)
    p0 = 0.0
    p1 = 0.0
    mask = 1 << (t - 1)
    for i in range(len(probs)):
        p = probs[i]
        if p > threshold:
            if (i & mask) == 0:
                p0 += p
            else:
                p1 += p
    return (float(p0), float(p1))


def measure_x(s, t, threshold=1e-10):
    """QSim.jl:320-324"""
    c = s.copy()
    u2(c, t, H)
    return measure_z(c, t, threshold)


def measure_y(s, t, threshold=1e-10):
    """QSim.jl:326-330"""
    c = s.copy()
    u2(c, t, rx_gate(np.pi / 2))
    return measure_z(c, t, threshold)


def statevector_to_density_matrix(s):
    """QSim.jl:39-41"""
    return np.outer(s, s.conj())


def partial_trace(rho, dim_A, dim_B, subsystem):
    """QInfo.jl:381-405, loop for loop."""
    assert rho.shape == (dim_A * dim_B, dim_A * dim_B), "Dimension mismatch"
    rho = (rho + rho.conj().T) / 2
    if subsystem == 1:
        out = np.zeros((dim_B, dim_B), dtype=np.complex128)
        for i in range(dim_B):
            for j in range(dim_B):
                for k in range(dim_A):
                    out[i, j] += rho[k * dim_B + i, k * dim_B + j]
        return out
    out = np.zeros((dim_A, dim_A), dtype=np.complex128)
    for i in range(dim_A):
        for j in range(dim_A):
            for l in range(dim_B):
                out[i, j] += rho[i * dim_B + l, j * dim_B + l]
    return out


def matrix_log2(rho):
    """QInfo.jl:24-28"""
    w, v = np.linalg.eigh(rho)
    w = np.maximum(w, 1e-10)
    return v @ np.diag(np.log2(w)) @ v.conj().T


def von_neumann_entropy(rho):
    """QInfo.jl:47-50"""
    rho = (rho + rho.conj().T) / 2
    return float(-np.real(np.trace(rho @ matrix_log2(rho))))


def entanglement_entropy(rho_AB, dim_A, dim_B):
    """QInfo.jl:178-182"""
    rho_AB = (rho_AB + rho_AB.conj().T) / 2
    return von_neumann_entropy(partial_trace(rho_AB, dim_A, dim_B, 2))


# ---- the Matrix{ComplexF64} (density matrix) methods, literally: op * rho * op' --------------------

def density_matrix(n, m):
    """density_matrix(n, m), QSim.jl:43-46: s * s' of the basis state"""
    s = statevector(n, m)
    return np.outer(s, np.conj(s))


def _op_1q(q, t, U):
    """kron(id(r), kron(U, id(l))), QSim.jl:71-73"""
    return np.kron(id_(q - t), np.kron(np.asarray(U, dtype=np.complex128), id_(t - 1)))


def u2_dm(rho, t, U):
    """u2!(rho, t, U), QSim.jl:67-79"""
    q = int(round(np.log2(rho.shape[0])))
    if not t <= q:
        raise RuntimeError("Qubit index out of range")
    op = _op_1q(q, t, U)
    return op @ rho @ op.conj().T


def controlled_dm(rho, c, t, V):
    """controlled!(rho, c, t, V), QSim.jl:209-232 — the 4x4 block sits directly above `mid`, as in the vector method"""
    q = int(round(np.log2(rho.shape[0])))
    if not (c <= q and t <= q):
        raise RuntimeError("Qubit index out of range")
    if c == t:
        raise RuntimeError("Control and target cannot be the same qubit")
    a, b = min(c, t), max(c, t)
    left, right = id_(q - b), id_(a - 1)
    mid = id_(b - a - 1) if (b - a - 1) > 0 else np.ones((1, 1), dtype=np.complex128)
    V = np.asarray(V, dtype=np.complex128)
    P0 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
    P1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)
    if c < t:
        U2 = np.kron(id_(1), P0) + np.kron(V, P1)
    else:
        U2 = np.kron(P0, id_(1)) + np.kron(P1, V)
    op = np.kron(left, np.kron(U2, np.kron(mid, right)))
    return op @ rho @ op.conj().T


def swap_dm(rho, q1, q2):
    """swap!(rho, q1, q2), QSim.jl:277-290"""
    q = int(round(np.log2(rho.shape[0])))
    if not (q1 <= q and q2 <= q):
        raise RuntimeError("Qubit index out of range")
    if q1 == q2:
        return rho
    rho = controlled_dm(rho, q1, q2, X)
    rho = controlled_dm(rho, q2, q1, X)
    return controlled_dm(rho, q1, q2, X)


def mp_dm(rho):
    """mp(rho) = real.(diag(rho)), QSim.jl:295"""
    return np.real(np.diag(rho)).copy()


def measure_z_dm(rho, t, threshold=1e-10):
    """measure_z(rho, t), QSim.jl:332-352"""
    q = int(round(np.log2(rho.shape[0])))
    if not t <= q:
        raise RuntimeError("Qubit index out of range")
    probs = np.real(np.diag(rho))
    p0 = p1 = 0.0
    mask = 1 << (t - 1)
    for i in range(len(probs)):
        if probs[i] > threshold:
            if (i & mask) == 0:
                p0 += probs[i]
            else:
                p1 += probs[i]
    return (p0, p1)


def measure_x_dm(rho, t, threshold=1e-10):
    """measure_x(rho, t), QSim.jl:354-358"""
    return measure_z_dm(u2_dm(rho.copy(), t, H), t, threshold)


def measure_y_dm(rho, t, threshold=1e-10):
    """measure_y(rho, t), QSim.jl:360-364"""
    return measure_z_dm(u2_dm(rho.copy(), t, rx_gate(np.pi / 2)), t, threshold)
