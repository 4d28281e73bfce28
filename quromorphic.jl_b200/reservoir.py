"""reservoir.py — the callers on either side of the hot path.

* `LSMNet`: host mirror of /root/reference/src/Neuromorphic/LSM.jl — the classical leaky-tanh reservoir
  that drives config C5.  Weights are injected (the reference draws them from Julia's MersenneTwister,
  LSM.jl:23-27, which cannot be reproduced outside Julia; the reference's own tests inject them the
  same way, test_LSM.jl:444-446).  50 x 50 real algebra: it stays on the host, as in the reference.
* `QuantumReservoir`: a quantum reservoir whose state lives on the B200 engine: per time step an
  encoding layer ry(pi * x_q), a fixed reservoir layer (rx, rz per qubit, CNOT chain) and the readouts
  the north star names — all-qubit <Z>, and the QInfo reduced-density / entropy readouts.
  `collect_states` / `train_readout` / `predict` mirror LSM.jl:49-72 with the quantum features.
"""
import numpy as np

from . import circuits as C
from . import qinfo as QI
from . import qsim as Q


class LSMNet:
    """LSM.jl:7-19 with injected weights (the positional constructor, LSM.jl:7-19)."""

    def __init__(self, Win, Wres, leak_rate=0.3, output_dim=1):
        self.Win = np.asarray(Win, dtype=np.float64)
        self.Wres = np.asarray(Wres, dtype=np.float64)
        self.reservoir_size, self.input_dim = self.Win.shape
        self.leak_rate = float(leak_rate)
        self.Wout = np.zeros((output_dim, self.reservoir_size))
        self.bout = np.zeros(output_dim)
        self.state = np.zeros(self.reservoir_size)

    def reset(self):
        """reset!, LSM.jl:39-41"""
        self.state[:] = 0.0

    def step(self, u):
        """step!, LSM.jl:43-47: x <- (1 - a) x + a tanh(Win u + Wres x)"""
        pre = self.Win @ np.asarray(u, dtype=np.float64) + self.Wres @ self.state
        self.state = (1 - self.leak_rate) * self.state + self.leak_rate * np.tanh(pre)
        return self.state

    def collect_states(self, inputs):
        """collect_states, LSM.jl:49-57 (inputs: input_dim x T)"""
        self.reset()
        T = inputs.shape[1]
        out = np.zeros((self.reservoir_size, T))
        for t in range(T):
            out[:, t] = self.step(inputs[:, t])
        return out


def ridge_readout(states, targets, reg=1e-6):
    """train_readout!, LSM.jl:59-67: (X X' + reg I) \\ X Y' with a bias row"""
    X = np.vstack([states, np.ones((1, states.shape[1]))])
    W = np.linalg.solve(X @ X.T + reg * np.eye(X.shape[0]), X @ targets.T)
    return W[:-1, :].T, W[-1, :]


class QuantumReservoir:
    def __init__(self, n, layer_seed=1203, mode="reference", device=0, entropy_k=0):
        self.n, self.mode, self.entropy_k = n, mode, entropy_k
        self.layer = C.reservoir_layer_params(n, layer_seed)
        self.state = Q.B200State(n, 0, device=device, mode=mode)
        self.Wout, self.bout = None, None

    def reset(self):
        self.state.reset(0)

    def step_gates(self, x):
        """one time step as a wire-format gate list: ry(pi x_q) encoding, reservoir layer, CNOT chain"""
        gl = C.GateList(self.n, self.mode)
        for q in range(1, self.n + 1):
            gl.ry(q, np.pi * float(x[(q - 1) % len(x)]))
        for q in range(1, self.n + 1):
            gl.rx(q, self.layer[q - 1][0])
            gl.rz(q, self.layer[q - 1][1])
        for q in range(1, self.n):
            gl.cnot(q, q + 1)
        return gl.array()

    def step(self, x):
        """evolve one step; returns the feature vector z_q = p0 - p1 (and the two entropies if asked)"""
        # one call per step: circuit + all-qubit <Z>; on a large register the sums are taken inside the last fused pass (QM_OPT_ZFOLD)
        z = self.state.apply_circuit_expect_z(self.step_gates(x))
        feats = z[:, 0] - z[:, 1]
        if self.entropy_k:
            k, n = self.entropy_k, self.n
            rho_B = QI.partial_trace(self.state, 1 << (n - k), 1 << k, 1)      # keep the low k qubits
            rho_A = QI.partial_trace(self.state, 1 << k, 1 << (n - k), 2)      # keep the high k qubits
            feats = np.concatenate([feats, [QI.von_neumann_entropy(rho_B), QI.von_neumann_entropy(rho_A)]])
        return feats

    def collect_states(self, drive):
        """drive: (dim, T) array; column t is the encoding vector of step t (e.g. LSMNet states)"""
        self.reset()
        return np.stack([self.step(drive[:, t]) for t in range(drive.shape[1])], axis=1)

    def train_readout(self, drive, targets, reg=1e-6):
        self.Wout, self.bout = ridge_readout(self.collect_states(drive), targets, reg)

    def predict(self, drive):
        """predict, LSM.jl:69-72"""
        return self.Wout @ self.collect_states(drive) + self.bout[:, None]
