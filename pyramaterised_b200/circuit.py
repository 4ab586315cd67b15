"""PQC -- the reference's circuit class (/root/reference/pyramaterised/circuit.py) on the
B200 engine.  Same attributes and methods; `run` lowers the flat gate list to one gate
program and launches it, instead of G sparse mat-vecs with per-parameter operator rebuilds.
Additive batch entry points (`run_batch`, `qfim_batch`) expose the sample axis the
reference loops over in Python (measure.py:132,247,356,407).
"""
import numpy as np
import torch

from . import engine
from .gates import *            # noqa: F401,F403  (the reference re-exports gates through circuit)
from .gates import rng
from .qobj import PauliSum, State, basis_state


class PQC():
    """An n qubit wide, p layer deep parameterised quantum circuit."""

    def __init__(self, n_qubits):
        self.n_qubits = n_qubits
        self.n_layers = 0
        self.layers = []
        self.gates = []
        self.parameterised = []
        self.n_params = 0
        self._program = None
        if n_qubits >= 2:
            self.set_H('ZZ')
        self._initial_state = None      # lazily |0..0> so that building a circuit needs no GPU
        self._init_index = 0
        self._state = None

    # initial_state / state mirror circuit.py:22-23 but are created on first use
    @property
    def initial_state(self):
        if self._initial_state is None:
            self._initial_state = basis_state(self.n_qubits, self._init_index)
        return self._initial_state

    def set_initial_basis_state(self, index):
        """Start from the computational basis state |index> (qubit 0 = most significant bit)."""
        self._init_index = int(index)
        self._initial_state = None

    @initial_state.setter
    def initial_state(self, st):
        self._initial_state = st if isinstance(st, State) else State(st)

    @property
    def state(self):
        return self.initial_state if self._state is None else self._state

    @state.setter
    def state(self, st):
        self._state = st

    def set_H(self, H):
        """'ZZ' = Z0 Z1 (circuit.py:25-33) or any PauliSum."""
        if isinstance(H, str) and H == 'ZZ':
            Z0 = genFockOp(qt.sigmaz(), 0, self.n_qubits, 2)
            Z1 = genFockOp(qt.sigmaz(), 1, self.n_qubits, 2)
            self.H = Z0 * Z1
        else:
            self.H = H

    def set_initial_state(self, state):
        """Every qubit in the same one-qubit state (circuit.py:35-36)."""
        self.initial_state = qt.tensor([state for i in range(self.n_qubits)])

    def add_layer(self, layer, n=1):
        for i in range(n):
            self.layers.append(deepcopy(layer))
        self.n_layers += n
        self.set_gates()

    def set_layer(self, layer, pos):
        self.layers[pos] = deepcopy(layer)
        self.set_gates()

    def get_layer(self, pos):
        return self.layers[pos]

    def set_gates(self):
        """Flatten layers into `gates`; `parameterised` / `n_params` keep the reference's
        bookkeeping, including its for...else that appends a -1 after EVERY gate and
        double counts parameters (quirk Q1, circuit.py:62-72)."""
        self.gates = [g for layer in self.layers for g in layer]
        self.parameterised = []
        total = 0
        for gate in self.gates:
            total += gate.param_count
            for _ in range(gate.param_count):
                total += 1
                self.parameterised.append(total)
            self.parameterised.append(-1)
        self.n_params = total
        self._program = None

    def get_params(self):
        angles = []
        for g in self.gates:
            if g.param_count == 2:
                angles += [g.theta, g.phi]
            elif g.param_count == 1:
                angles.append(g.theta)
        return angles

    # ---- lowering -----------------------------------------------------------------------
    @property
    def n_true_params(self):
        return sum(g.param_count for g in self.gates)

    def lower(self):
        """-> list of primitive ops (kind, q0, q1, param, param2, group, scale, offset)."""
        ops, slot = [], 0
        for gi, g in enumerate(self.gates):
            for o in g._lower(slot if g.param_count > 0 else -1):
                ops.append(o[:5] + (gi,) + o[6:])
            slot += g.param_count
        return ops

    @property
    def program(self):
        if self._program is None:
            if any(hasattr(g, "_lam") for g in self.gates):        # dense gates (ARBGATE)
                self._program = engine.SegmentedProgram(self.n_qubits, self.gates)
            else:
                self._program = engine.Program(self.n_qubits, self.n_true_params, self.lower())
        return self._program

    def _derivatives_exact(self):
        for g in self.gates:
            chk = getattr(g, "_sum_of_generators_is_exact", None)
            if chk is not None and not chk():
                return False
        return True

    # ---- parameters (circuit.py:86-116) ------------------------------------------------------
    def set_params(self, angles):
        """Angles from a list in gate order, or "random" draws from the module RNG."""
        raw = []
        k = 0
        for g in (g for g in self.gates if g.is_param):
            if g.param_count == 2:
                if type(angles) != str:
                    a1, a2 = angles[k], angles[k + 1]
                elif angles == "random":
                    a1 = rng.random(1)[0] * 2 * np.pi
                    a2 = rng.random(1)[0] * 2 * np.pi
                else:
                    raise Exception("No parameters supplied!")
                g.set_theta(a1)
                g.set_phi(a2)
                raw += [a1, a2]
                k += 2
            elif g.param_count == 1:
                if type(angles) != str:
                    a1 = angles[k]
                elif angles == "random":
                    a1 = rng.random(1)[0] * 2 * np.pi
                else:
                    raise Exception("No parameters supplied!")
                g.set_theta(a1)
                raw.append(a1)
                k += 1
        self._raw_angles = np.array(raw, dtype=np.float64)
        return self._raw_angles

    def draw_random(self, S):
        """The S x P angles S successive run("random") calls would draw (gates.py:10,
        circuit.py:100-101,112)."""
        P = self.n_true_params
        return (rng.random(S * P) * 2 * np.pi).reshape(S, P)

    # ---- simulation -----------------------------------------------------------------------------
    def run(self, angles):
        """|psi> = U_G ... U_1 |init> (circuit.py:118-125)."""
        raw = self.set_params(angles)
        out = self.program.run(raw.reshape(1, -1) if len(raw) else None,
                               init=self.initial_state.tensor)
        return State(out[0], self.initial_state.dims)

    def run_batch(self, angles, S=None):
        """Batched run: angles [S,P] array (or "random" with S) -> device tensor [S, 2^n].
        Equivalent to S calls of run(); gates keep the last row's angles like the reference."""
        if isinstance(angles, str):
            if angles != "random":
                raise Exception("No parameters supplied!")
            angles = self.draw_random(S)
        P = self.n_true_params
        if P == 0:
            n = S if S is not None else (len(angles) if hasattr(angles, "__len__") else 1)
            return self.program.run(None, init=self.initial_state.tensor).expand(n, -1).contiguous()
        out = self.program.run(angles, init=self.initial_state.tensor)
        if not hasattr(angles, "is_cuda") and len(angles):
            self.set_params(list(np.asarray(angles)[-1]))
        return out

    def update_state(self, angles):
        self.state = self.run(angles)
        return self.state

    def cost(self, angles):
        """<psi|H|psi> (circuit.py:132-137)."""
        self.state = self.run(angles=angles)
        return qt.expect(self.H, self.state)

    def fidelity(self, target_state):
        return np.abs(self.state.overlap(target_state)) ** 2

    def flip_deriv(self):
        for g in (g for g in self.gates if g.is_param):
            g.flip_pauli()

    # ---- derivative states (circuit.py:149-192) ----------------------------------------------------
    def _gradient_buffer_literal(self):
        """circuit.py:149-192 followed gate by gate: for every parameterised gate the circuit is
        re-run with that gate replaced by `derivative() * gate`.  The slow path, taken only when a
        gate's derivative is not a sum of Pauli generators the engine can spawn (a
        shared_parameter block whose members do not commute, gates.py:458-466): P x G small
        gate-program launches instead of one derivative pipeline."""
        if any(g.param_count == 2 for g in self.gates):
            raise NotImplementedError("two-parameter gates together with a non-commuting "
                                      "shared_parameter block are not lowered")
        rows = [self.initial_state]
        for g in self.gates:
            rows[0] = g * rows[0]
        for g_on in (g for g in self.gates if g.param_count > 0):
            deriv = g_on.derivative()
            st = self.initial_state
            for g in self.gates:
                st = g * st
                if g is g_on:
                    st = deriv * st                 # (deriv * gate) * state, circuit.py:164-168
            rows.append(st)
        return torch.stack([r.tensor.reshape(-1) for r in rows])

    def _gradient_buffer(self):
        if self.program.grad_supported and not self._derivatives_exact():
            return self._gradient_buffer_literal()
        if not self.program.grad_supported:
            raise NotImplementedError("derivative states for this gate set are not lowered yet")
        # quirk Q2 (circuit.py:186-189): for a two-parameter gate the reference re-finds the
        # gate with an index into the PARAMETERISED list; that is the intended gate only when
        # every earlier gate is parameterised.  Anything else is fenced off, not imitated.
        count = 0
        for loc, g in enumerate(self.gates):
            if g.param_count == 0:
                continue
            if g.param_count == 2 and count != loc:
                raise NotImplementedError("two-parameter gate after a non-parameterised gate: the "
                                          "reference differentiates the wrong gate here (quirk Q2)")
            count += 1
        raw = np.array([a for g in self.gates if g.param_count > 0
                        for a in self._raw_of(g)], dtype=np.float64)
        return self.program.gradients(raw.reshape(1, -1), init=self.initial_state.tensor)[0]

    @staticmethod
    def _raw_of(g):
        """The un-transformed parameter(s) a gate was last set with."""
        if g.param_count == 2:
            return [g.theta, g.phi]
        th = g.theta
        if type(g).__name__ == "negative_R_z":
            th = -th
        elif type(g).__name__ == "offset_R_z":
            th = th - g.offset
        return [th]

    def take_derivative(self, g_on, param=0):
        """Derivative state w.r.t. the parameter of gate `g_on` (identity lookup, like
        self.gates.index in the reference, circuit.py:156)."""
        g_loc = next(i for i, g in enumerate(self.gates) if g is g_on)
        slot = sum(g.param_count for g in self.gates[:g_loc])
        if g_on.param_count == 0:
            return self.run(self.get_params_raw())
        buf = self._gradient_buffer()
        return State(buf[1 + slot + max(0, param - 1)], self.initial_state.dims)

    def get_params_raw(self):
        return [a for g in self.gates if g.param_count > 0 for a in self._raw_of(g)]

    def get_gradients(self):
        """One derivative state per parameter, in parameter order (circuit.py:174-192)."""
        buf = self._gradient_buffer()
        dims = self.initial_state.dims
        out = [State(buf[1 + p], dims) for p in range(self.n_true_params)]
        self._last_gradient_buffer = buf
        return out

    # ---- fused batch QFIM (additive) -----------------------------------------------------------------
    def qfim_batch(self, angles, want_states=False, max_work_bytes=None):
        """angles [S,P] -> QFIM [S,P,P] (device), as update_state + get_QFI per row."""
        if not self.program.grad_supported:
            raise NotImplementedError("QFIM for this gate set is not lowered yet")
        if not self._derivatives_exact():
            # literal derivative states row by row (see _gradient_buffer_literal)
            from . import engine
            a = np.asarray(angles.cpu() if hasattr(angles, "cpu") else angles, dtype=np.float64)
            Fs, sts = [], []
            for row in a.reshape(-1, self.n_true_params):
                self.set_params(list(row))
                buf = self._gradient_buffer_literal()
                Fs.append(engine.qfim_from_grads(buf[:1], buf[1:].unsqueeze(0))[0])
                sts.append(buf[0])
            F = torch.stack(Fs)
            return (F, torch.stack(sts)) if want_states else F
        return self.program.qfim(angles, init=self.initial_state.tensor, want_states=want_states,
                                 max_work_bytes=max_work_bytes)

    def __repr__(self):
        line = f"A {self.n_qubits} qubit, {self.n_layers} layer deep PQC. \n"
        for count, l in enumerate(self.layers):
            line += f"Layer {count}: {l} \n"
        return line
