"""Device-side plumbing between the reference-facing API and the C ABI.

PyTorch is used for exactly three things here: allocating device buffers, naming the
current CUDA stream, and host<->device copies.  All arithmetic happens in
libpqc_b200.so (csrc/*.cu).  No function in this module has a CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_device_ok = False
QFIM_WORK_BYTES = 32 << 30   # default cap of the live-vector workspace of Program.qfim
gpu_launches = 0          # kernels enqueued through this module (bench.py reports it)


def device():
    """The CUDA device everything runs on; raises when there is none (no CPU fallback)."""
    global _device_ok
    if not torch.cuda.is_available():
        raise RuntimeError("pyramaterised_b200 needs a CUDA device (built for B200, sm_100a); "
                           "there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    if not _device_ok:
        lib = _lib.load()
        torch.cuda.init()
        torch.zeros(1, device=dev)                 # make sure the primary context exists
        maj, mnr, sms = C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.pqc_device_check(C.byref(maj), C.byref(mnr), C.byref(sms)))
        _device_ok = True
    return dev


def launch_count():
    """Kernels launched by libpqc_b200.so in this process (counted inside the library)."""
    return int(_lib.load().pqc_launch_count())


PASS_KERNELS = ("k_apply_pass", "k_sweep_pass", "k_layer_pass", "k_layer_seq", "k_tile_pipe")


def profile_begin():
    _lib.check(_lib.load().pqc_profile_begin())


def profile_end():
    """-> dict(ms, launches, bytes) for the gate-apply kernel since profile_begin()."""
    out = (C.c_double * 4)()
    _lib.check(_lib.load().pqc_profile_end(out))
    kinds = (C.c_double * (3 * len(PASS_KERNELS)))()
    _lib.check(_lib.load().pqc_profile_kinds(kinds, len(PASS_KERNELS)))
    by = {name: {"ms": kinds[3 * k], "launches": int(kinds[3 * k + 1]), "bytes": kinds[3 * k + 2]}
          for k, name in enumerate(PASS_KERNELS) if kinds[3 * k + 1] > 0}
    return {"ms": out[0], "launches": int(out[1]), "bytes": out[2], "by_kernel": by}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _count(k=1):
    global gpu_launches
    gpu_launches += k


def as_states(x, dev=None):
    """[S, D] (or [D]) complex128 contiguous device tensor from numpy / torch input."""
    dev = dev or device()
    if isinstance(x, torch.Tensor):
        t = x.to(device=dev, dtype=torch.complex128)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))).to(dev)
    return t.contiguous()


def as_angles(angles, P, dev=None):
    """[S, >=P] float64 contiguous device tensor.  Pinned host tensors copy asynchronously."""
    dev = dev or device()
    if isinstance(angles, torch.Tensor):
        t = angles
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        if t.device != dev:
            t = t.to(dev, non_blocking=True)
    else:
        a = np.asarray(angles, dtype=np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    if t.dim() == 1:
        t = t.reshape(1, -1)
    if t.shape[1] < P:
        raise IndexError("list index out of range")       # circuit.py:97,110 on a short list
    return t.contiguous()


class Program:
    """A lowered gate program (pqc_program handle).  ops: sequence of
    (kind, q0, q1, param, param2, group, scale, offset).

    A Program is bound to the CUDA device of its first launch (the library refuses another
    one) and keeps per-program scratch -- the per-sample trig table, the QFIM workspace -- so it
    is NOT thread-safe and must not run on two streams at once: callers that overlap work
    (dist.streamed_expressibility's side stream) order the launches of one Program with stream
    waits.  release_workspace() frees the QFIM workspace (up to QFIM_WORK_BYTES)."""

    def __init__(self, n_qubits, n_params, ops):
        lib = _lib.load()               # planning is host-only; the device is needed to run
        self.n = int(n_qubits)
        self.P = int(n_params)
        self.ops = list(ops)
        arr = (_lib.PqcOp * max(1, len(self.ops)))()
        for i, o in enumerate(self.ops):
            arr[i] = _lib.PqcOp(*[int(v) for v in o[:6]], float(o[6]), float(o[7]))
        h = C.c_void_p()
        _lib.check(lib.pqc_program_create(self.n, self.P, len(self.ops), arr, C.byref(h)))
        self._h = h
        st = (C.c_int64 * 8)()
        _lib.check(lib.pqc_program_stats(self._h, st))
        self.n_passes = int(st[3])
        self.tile_bits = int(st[4])
        self.grad_supported = bool(st[5])
        self.n_qfim_passes = int(st[6])

    @property
    def dim(self):
        return 1 << self.n

    def release_workspace(self):
        """Free the live-vector workspace Program.qfim keeps between calls."""
        self._work = None

    def describe(self):
        """The execution plan, one line per stage (host-only)."""
        buf = C.create_string_buffer(1 << 20)
        _lib.check(_lib.load().pqc_program_describe(self._h, buf, len(buf)))
        return buf.value.decode()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().pqc_program_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- PQC.run (circuit.py:118-125) ---------------------------------------------------
    def run(self, angles, init=None, out=None):
        """angles [S,P] -> states [S,D].  init: None (|0..0>), [D] (shared) or [S,D]."""
        dev = device()
        a = as_angles(angles, self.P, dev) if self.P > 0 else None
        S = a.shape[0] if a is not None else (init.shape[0] if init is not None and init.dim() == 2 else 1)
        D = self.dim
        if out is None:
            out = torch.empty((S, D), dtype=torch.complex128, device=dev)
        if S == 0:
            return out
        stride = 0
        if init is not None:
            init = as_states(init, dev)
            if init.dim() == 2:
                if init.shape[0] != S:
                    raise ValueError("per-sample initial states must match the angle batch")
                stride = D
        _lib.check(_lib.load().pqc_run_batch(self._h, _p(a), a.shape[1] if a is not None else 0, S,
                                             _p(init), stride, _p(out), _stream()))
        _count(self.n_passes)
        return out

    # ---- PQC.get_gradients (circuit.py:149-192) -------------------------------------------
    def gradients(self, angles, init=None):
        """-> buffer [S, P+1, D]: [:,0] final state, [:,1+p] derivative state p."""
        dev = device()
        a = as_angles(angles, self.P, dev) if self.P > 0 else None
        S = a.shape[0] if a is not None else 1
        buf = torch.empty((S, self.P + 1, self.dim), dtype=torch.complex128, device=dev)
        stride = 0
        if init is not None:
            init = as_states(init, dev)
            if init.dim() == 2:
                if init.shape[0] != S:
                    raise ValueError("per-sample initial states must match the angle batch")
                stride = self.dim
        _lib.check(_lib.load().pqc_gradients_batch(self._h, _p(a), a.shape[1] if a is not None else 0,
                                                   S, _p(init), stride, _p(buf), _stream()))
        _count(self.n_qfim_passes + self.P)
        return buf

    # ---- update_state + get_QFI fused over a batch ---------------------------------------
    def qfim(self, angles, init=None, want_states=False, max_work_bytes=None):
        dev = device()
        lib = _lib.load()
        a = as_angles(angles, self.P, dev)
        S = a.shape[0]
        if S == 0:
            F = torch.empty((0, self.P, self.P), dtype=torch.float64, device=dev)
            st = torch.empty((0, self.dim), dtype=torch.complex128, device=dev)
            return (F, st) if want_states else F
        need = C.c_int64()
        _lib.check(lib.pqc_qfim_workspace_bytes(self._h, S, C.byref(need)))
        per = (need.value - 256) // max(1, S)
        if max_work_bytes is None:
            max_work_bytes = QFIM_WORK_BYTES
        nbytes = min(need.value, max(per + 256, (max_work_bytes // per) * per + 256))
        work = getattr(self, "_work", None)
        if work is None or work.numel() < nbytes or work.device != dev:
            self._work = work = None           # drop the old block before asking for a bigger one
            self._work = work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        nbytes = work.numel()
        F = torch.empty((S, self.P, self.P), dtype=torch.float64, device=dev)
        states = torch.empty((S, self.dim), dtype=torch.complex128, device=dev) if want_states else None
        if init is not None:
            init = as_states(init, dev)
        _lib.check(lib.pqc_qfim_batch(self._h, _p(a), a.shape[1], S, _p(init), _p(work), nbytes,
                                      _p(F), _p(states), _stream()))
        _count(self.n_qfim_passes + 3 * self.P + 1)
        return (F, states) if want_states else F


def dense_apply(states, M, lam=None, theta=None, deriv=False, out=None):
    """out[s] = M (w_s * states[s]) with w_s = 1, or exp(-i theta_s lam) (times -i lam / 2 with
    `deriv`): the two halves of an ARBGATE (gates.py:407-435) in its eigenbasis."""
    states = states.contiguous()
    S, D = states.shape
    if out is None:
        out = torch.empty_like(states)
    if S == 0:
        return out
    if theta is not None:
        theta = theta.contiguous()
    _lib.check(_lib.load().pqc_dense_apply_batch(_p(states), S, D.bit_length() - 1, _p(M), _p(lam),
                                                 _p(theta), 1 if theta is not None else 0,
                                                 1 if deriv else 0, _p(out), _stream()))
    _count()
    return out


class SegmentedProgram:
    """A circuit that contains dense gates (ARBGATE): gate-program segments (Program) joined by
    dense eigenbasis products.  Same surface as Program (run / gradients / qfim / describe), so
    PQC and Measurements do not care.  Derivative states follow circuit.py:149-192: the vector of
    parameter p is created inside its segment (or as (-i H / 2) exp(-i theta H) psi for a dense
    gate) and carried through every later segment."""

    def __init__(self, n_qubits, gates):
        self.n = int(n_qubits)
        self.segs = []                  # ("ops", Program, lo, hi) | ("dense", gate, slot)
        ops, lo, slot = [], 0, 0
        for gi, g in enumerate(gates):
            if hasattr(g, "_lam"):
                if g.q_N != self.n:
                    raise ValueError("ARBGATE Hamiltonian does not match the register size")
                if ops:
                    self.segs.append(("ops", Program(self.n, slot - lo, ops), lo, slot))
                self.segs.append(("dense", g, slot))
                ops, slot = [], slot + 1
                lo = slot
                continue
            for o in g._lower(slot - lo if g.param_count > 0 else -1):
                ops.append(o[:5] + (gi,) + o[6:])
            slot += g.param_count
        if ops:
            self.segs.append(("ops", Program(self.n, slot - lo, ops), lo, slot))
        self.P = slot
        self.grad_supported = all(s[1].grad_supported for s in self.segs if s[0] == "ops")
        self.n_passes = sum(s[1].n_passes if s[0] == "ops" else 2 for s in self.segs)
        self.n_qfim_passes = sum(s[1].n_qfim_passes if s[0] == "ops" else 2 for s in self.segs)
        self._dev = {}

    @property
    def dim(self):
        return 1 << self.n

    def describe(self):
        lines = []
        for s in self.segs:
            if s[0] == "ops":
                lines.append(f"SEGMENT gate program, parameters [{s[2]}, {s[3]})")
                lines.append(s[1].describe())
            else:
                lines.append(f"SEGMENT dense gate {s[1]!r}, parameter {s[2]}: V^dagger, phase, V")
        return "\n".join(lines)

    def _eig(self, gate):
        key = id(gate)
        if key not in self._dev:
            dev = device()
            V = torch.from_numpy(np.ascontiguousarray(gate._V)).to(dev)
            Vh = torch.from_numpy(np.ascontiguousarray(gate._V.conj().T)).to(dev)
            lam = torch.from_numpy(np.ascontiguousarray(gate._lam)).to(dev)
            self._dev[key] = (V, Vh, lam)
        return self._dev[key]

    def _dense(self, gate, vecs, theta, deriv=False):
        V, Vh, lam = self._eig(gate)
        y = dense_apply(vecs, Vh)
        return dense_apply(y, V, lam, theta, deriv)

    def _seg_run(self, seg, a, vecs, rep=1):
        """One segment applied to vecs [S * rep, D] (sample-major: row s * rep + r)."""
        if seg[0] == "ops":
            prog, lo, hi = seg[1], seg[2], seg[3]
            ang = a[:, lo:hi].repeat_interleave(rep, dim=0) if hi > lo and rep > 1 else a[:, lo:hi]
            return prog.run(ang.contiguous() if hi > lo else None, init=vecs)
        th = a[:, seg[2]]
        return self._dense(seg[1], vecs, th.repeat_interleave(rep) if rep > 1 else th)

    def _start(self, a, init, S):
        dev = device()
        if init is None:
            v = torch.zeros((S, self.dim), dtype=torch.complex128, device=dev)
            v[:, 0] = 1.0
            return v
        init = as_states(init, dev)
        return init.reshape(1, -1).expand(S, -1).contiguous() if init.dim() == 1 else init

    def run(self, angles, init=None, out=None):
        dev = device()
        if self.P > 0:
            a = as_angles(angles, self.P, dev)
            S = a.shape[0]
        else:
            a = torch.empty((1, 0), dtype=torch.float64, device=dev)
            S = init.shape[0] if init is not None and hasattr(init, "dim") and init.dim() == 2 else 1
        v = self._start(a, init, S)
        for seg in self.segs:
            v = self._seg_run(seg, a, v)
        if out is not None:
            out.copy_(v)
            return out
        return v

    def gradients(self, angles, init=None):
        """-> [S, P+1, D]: [:, 0] the final state, [:, 1 + p] derivative state p."""
        dev = device()
        a = as_angles(angles, self.P, dev)
        S, D = a.shape[0], self.dim
        buf = torch.empty((S, self.P + 1, D), dtype=torch.complex128, device=dev)
        psi = self._start(a, init, S)
        for k, seg in enumerate(self.segs):
            if seg[0] == "ops":
                prog, lo, hi = seg[1], seg[2], seg[3]
                if hi > lo:
                    g = prog.gradients(a[:, lo:hi].contiguous(), init=psi)     # [S, Pk + 1, D]
                    new, psi = g[:, 1:], g[:, 0].contiguous()
                else:
                    new, psi = None, prog.run(None, init=psi)
            else:
                lo, hi = seg[2], seg[2] + 1
                th = a[:, lo].contiguous()
                new = self._dense(seg[1], psi, th, deriv=True).unsqueeze(1)
                psi = self._dense(seg[1], psi, th)
            if new is None:
                continue
            m = hi - lo
            vecs = new.reshape(S * m, D).contiguous()
            for later in self.segs[k + 1:]:
                vecs = self._seg_run(later, a, vecs, rep=m)
            buf[:, 1 + lo:1 + hi] = vecs.reshape(S, m, D)
        buf[:, 0] = psi
        return buf

    def qfim(self, angles, init=None, want_states=False, max_work_bytes=None):
        dev = device()
        a = as_angles(angles, self.P, dev)
        S = a.shape[0]
        F = torch.empty((S, self.P, self.P), dtype=torch.float64, device=dev)
        st = torch.empty((S, self.dim), dtype=torch.complex128, device=dev) if want_states else None
        per = (self.P + 1) * self.dim * 16 * 3
        chunk = max(1, int((max_work_bytes or QFIM_WORK_BYTES) // per))
        for c0 in range(0, S, chunk):
            sl = slice(c0, min(S, c0 + chunk))
            ini = init[sl] if init is not None and hasattr(init, "dim") and init.dim() == 2 else init
            g = self.gradients(a[sl], init=ini)
            F[sl] = qfim_from_grads(g[:, 0].contiguous(), g[:, 1:].contiguous())
            if want_states:
                st[sl] = g[:, 0]
        return (F, st) if want_states else F


# ---------------------------------------------------------------------------------------
# measures
# ---------------------------------------------------------------------------------------
def qfim_from_grads(states, grads):
    """states [S,D], grads [S,P,D] -> F [S,P,P] (measure.py:33-71)."""
    S, P, D = grads.shape
    F = torch.empty((S, P, P), dtype=torch.float64, device=grads.device)
    n = D.bit_length() - 1
    _lib.check(_lib.load().pqc_qfim_from_grads(_p(states), _p(grads), n, P, S, _p(F), _stream()))
    _count(2)
    return F


def eigvalsh(mats):
    """[S,P,P] float64 -> ascending eigenvalues [S,P] (scipy.linalg.eigh, measure.py:74,84)."""
    mats = mats.contiguous()
    S, P, _ = mats.shape
    out = torch.empty((S, P), dtype=torch.float64, device=mats.device)
    _lib.check(_lib.load().pqc_eigvalsh_batch(_p(mats), S, P, _p(out), _stream()))
    _count()
    return out


def eigh(mats):
    """[S,P,P] -> (eigenvalues [S,P] ascending, eigenvectors [S,P,P] as columns)."""
    mats = mats.contiguous()
    S, P, _ = mats.shape
    w = torch.empty((S, P), dtype=torch.float64, device=mats.device)
    v = torch.empty((S, P, P), dtype=torch.float64, device=mats.device)
    _lib.check(_lib.load().pqc_eigh_batch(_p(mats), S, P, _p(w), _p(v), _stream()))
    _count()
    return w, v


def count_greater(vals, cutoff):
    vals = vals.contiguous()
    rows, cols = vals.shape
    out = torch.empty((rows,), dtype=torch.int32, device=vals.device)
    _lib.check(_lib.load().pqc_count_greater(_p(vals), rows, cols, float(cutoff), _p(out), _stream()))
    _count()
    return out


def meyer_wallach(states):
    """[S,D] -> Q[S] (measure.py:226-249)."""
    states = states.contiguous()
    S, D = states.shape
    out = torch.empty((S,), dtype=torch.float64, device=states.device)
    if S == 0:
        return out
    _lib.check(_lib.load().pqc_meyer_wallach(_p(states), S, D.bit_length() - 1, _p(out), _stream()))
    _count(2)
    return out


def ptrace_1q(state, qubit):
    state = state.contiguous()
    n = state.numel().bit_length() - 1
    if not 0 <= qubit < n:
        raise IndexError("Invalid selection index in ptrace.")
    out = torch.empty((2, 2), dtype=torch.complex128, device=state.device)
    _lib.check(_lib.load().pqc_ptrace_1q(_p(state), n, int(qubit), _p(out), _stream()))
    _count(2)
    return out


def overlap(a, b):
    """<a_i|b_i> for rows of a, b ([D] or [S,D]) -> complex tensor [S].  A single row on
    either side is broadcast against the other (stride 0)."""
    a2 = a.reshape(-1, a.shape[-1]).contiguous()
    b2 = b.reshape(-1, b.shape[-1]).contiguous()
    D = a2.shape[1]
    if b2.shape[1] != D or (a2.shape[0] != b2.shape[0] and 1 not in (a2.shape[0], b2.shape[0])):
        raise TypeError("Can only calculate overlap for state vector Qobjs")
    S = max(a2.shape[0], b2.shape[0])
    sa = D if a2.shape[0] == S else 0
    sb = D if b2.shape[0] == S else 0
    out = torch.empty((S,), dtype=torch.complex128, device=a2.device)
    _lib.check(_lib.load().pqc_overlap_batch(_p(a2), sa, _p(b2), sb, D, S, _p(out), _stream()))
    _count()
    return out


def n_bins(n_pairs):
    """bins=int((75/10000)*len(F_samples)) (measure.py:153-154)."""
    return int((75 / 10000) * n_pairs)


def fidelity_hist(A, B=None, bins=0, hist=None, want_F=False):
    """Pairwise fidelities of state blocks.  B None -> all unordered pairs of A in
    itertools.combinations order (measure.py:133-136).  Returns (hist or None, F or None)."""
    A = A.contiguous()
    SA, D = A.shape
    tri = B is None
    Bm = A if tri else B.contiguous()
    SB = Bm.shape[0]
    n = D.bit_length() - 1
    if bins > 0 and hist is None:
        hist = torch.zeros((bins,), dtype=torch.int64, device=A.device)
    F = None
    if want_F:
        F = torch.empty((SA * (SA - 1) // 2,) if tri else (SA, SB), dtype=torch.float64,
                        device=A.device)
    if SA > 0 and SB > 0:
        _lib.check(_lib.load().pqc_fidelity_hist(_p(A), SA, _p(Bm), SB, n, int(tri), int(bins),
                                                 _p(hist), _p(F), _stream()))
        _count()
    return hist, F


def hist_f64(F, bins):
    F = F.contiguous()
    if bins <= 0:
        raise ValueError("`bins` must be positive, when an integer")
    hist = torch.zeros((bins,), dtype=torch.int64, device=F.device)
    _lib.check(_lib.load().pqc_hist_f64(_p(F), F.numel(), int(bins), _p(hist), _stream()))
    _count()
    return hist


def kl_haar(hist, hilbert_dim):
    """Measurements.expr on histogram counts (measure.py:161-180) -> 0-d device tensor."""
    hist = hist.contiguous()
    out = torch.empty((1,), dtype=torch.float64, device=hist.device)
    scratch = torch.empty((4,), dtype=torch.float64, device=hist.device)
    _lib.check(_lib.load().pqc_kl_haar(_p(hist), hist.numel(), float(hilbert_dim), _p(out),
                                       _p(scratch), _stream()))
    _count(3)
    return out


def magic(states, alphas=(2.0,)):
    """[S,D] -> [len(alphas), S] Renyi stabilizer entropies (measure.py:318-349)."""
    states = states.contiguous()
    S, D = states.shape
    al = (C.c_double * len(alphas))(*[float(x) for x in alphas])
    out = torch.empty((len(alphas), S), dtype=torch.float64, device=states.device)
    if S == 0:
        return out
    _lib.check(_lib.load().pqc_magic_batch(_p(states), S, D.bit_length() - 1, len(alphas), al,
                                           _p(out), _stream()))
    _count(2)
    return out


def _terms(pauli_terms):
    arr = (_lib.PqcPauliTerm * max(1, len(pauli_terms)))()
    for i, (xm, zm, c) in enumerate(pauli_terms):
        arr[i] = _lib.PqcPauliTerm(int(xm), int(zm), float(np.real(c)), float(np.imag(c)))
    return arr


def pauli_expect(states, pauli_terms):
    """<psi_s|H|psi_s> for H = sum coef * Pauli(xmask, zmask) -> complex [S]."""
    states = states.contiguous()
    S, D = states.shape
    out = torch.empty((S,), dtype=torch.complex128, device=states.device)
    _lib.check(_lib.load().pqc_pauli_expect_batch(_p(states), S, D.bit_length() - 1,
                                                  len(pauli_terms), _terms(pauli_terms), _p(out),
                                                  _stream()))
    _count()
    return out


def pauli_apply(states, pauli_terms):
    states = states.contiguous()
    S, D = states.shape
    out = torch.empty_like(states)
    _lib.check(_lib.load().pqc_pauli_apply_batch(_p(states), S, D.bit_length() - 1,
                                                 len(pauli_terms), _terms(pauli_terms), _p(out),
                                                 _stream()))
    _count()
    return out
