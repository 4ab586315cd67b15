"""Measurements -- the reference's capacity-measure class
(/root/reference/pyramaterised/measure.py) running on the B200 engine.

Same method names, arguments and return types.  Where the reference loops over samples
or state pairs in Python with one QuTiP call each (measure.py:132-136,247-248,356-358,
405-447), the sample set is generated and measured as ONE batch on the device.
"""
import random
from itertools import combinations            # noqa: F401

import numpy as np
import scipy.optimize
import torch

from . import engine
from .qobj import State


def _as_state(x):
    return x if isinstance(x, State) else State(x)


def _stack(states):
    """list of State -> [S, D] device tensor; a zero-copy view when the states already sit
    back to back in one buffer (as PQC.get_gradients returns them)."""
    ts = [_as_state(s).tensor for s in states]
    if not ts:
        return None
    D = ts[0].numel()
    p0 = ts[0].data_ptr()
    if all(t.is_contiguous() and t.data_ptr() == p0 + i * D * 16 and
           t.untyped_storage().data_ptr() == ts[0].untyped_storage().data_ptr()
           for i, t in enumerate(ts)):
        return ts[0].as_strided((len(ts), D), (D, 1))
    return torch.stack(ts)


class Measurements:
    def __init__(self, QC):
        self.QC = QC
        try:
            self.minimize_function = QC.cost
        except AttributeError:
            self.minimize_function = lambda x: x

    def set_minimise_function(self, function):
        self.minimize_function = function

    # ---- sample generation ---------------------------------------------------------------
    def _random_states(self, sample_N):
        """sample_N x QC.run("random") (measure.py:132,247,356,407) as one device batch."""
        if sample_N <= 0:
            return None
        # stand-in circuits (tests.py:12-60) subclass PQC without calling __init__: they have
        # run() but no gate list, and must be asked BEFORE any angle is drawn from the module RNG
        if hasattr(self.QC, "run_batch") and hasattr(self.QC, "gates") and hasattr(self.QC, "_program"):
            return self.QC.run_batch("random", sample_N)
        return _stack([self.QC.run("random") for _ in range(sample_N)])

    # ---- QFIM (measure.py:33-99) -------------------------------------------------------------
    def get_QFI(self, grad_list=[]):
        """F_pq = 4 Re(<d_p|d_q> - conj<psi|d_p><psi|d_q>) as an n_param x n_param array."""
        n_params = len([i for i in self.QC.parameterised if i > -1])
        grad_state_list = self.QC.get_gradients() if grad_list == [] else grad_list
        if n_params == 0:
            return np.zeros([0, 0])
        grads = _stack(grad_state_list[:n_params]).unsqueeze(0)
        psi = _as_state(self.QC.state).tensor.reshape(1, -1)
        return engine.qfim_from_grads(psi, grads)[0].cpu().numpy()

    def get_eigenvalues(self, QFI):
        m = torch.as_tensor(np.asarray(QFI, dtype=np.float64), device=engine.device())[None]
        w, v = engine.eigh(m)
        return w[0].cpu().numpy(), v[0].cpu().numpy()

    def get_effective_quantum_dimension(self, cutoff_eigvals):
        QFI = self.get_QFI()
        if QFI.shape[0] == 0:
            return 0
        m = torch.as_tensor(QFI, device=engine.device())[None]
        w = engine.eigvalsh(m)
        return int(engine.count_greater(w, cutoff_eigvals)[0].item())

    def new_measure(self, QFI=None):
        if QFI is None:
            QFI = self.get_QFI()
        eigvals, _ = self.get_eigenvalues(QFI)
        return sum([1 if v > 1 else v for v in eigvals])

    def find_overparam_point(self, layer_index_list, epsilon=1e-3):
        layers_to_add = [self.QC.get_layer(i) for i in layer_index_list]
        prev_rank, rank_diff, count = 0, 1, 0
        while rank_diff > epsilon and count < 1e6:
            for l in layers_to_add:
                self.QC.add_layer(l)
            self.QC.update_state("random")
            QFI = self.get_QFI()
            rank = np.linalg.matrix_rank(QFI)
            rank_diff = np.abs(rank - prev_rank)
            print(f"Iteration {count}, r0={prev_rank}, r1={rank}, delta = {rank_diff}")
            prev_rank = rank
            count += 1
        return count

    # ---- expressibility (measure.py:123-197) --------------------------------------------------
    def _gen_f_samples(self, sample_N):
        """|<psi_i|psi_j>|^2 for all i<j in itertools.combinations order, as a list."""
        states = self._random_states(sample_N)
        if states is None or states.shape[0] < 2:
            return []
        _, F = engine.fidelity_hist(states, want_F=True)
        return F.cpu().numpy().tolist()

    def _gen_histo(self, F_samples, filt=0):
        """(prob, bin midpoints); `filt` does nothing, as in the reference, because the
        unfiltered samples are the ones histogrammed (measure.py:150-154)."""
        bins = engine.n_bins(len(F_samples))
        F = torch.as_tensor(np.asarray(F_samples, dtype=np.float64), device=engine.device())
        counts = engine.hist_f64(F, bins).cpu().numpy()
        prob = counts / sum(counts)
        edges = np.linspace(0, 1, bins + 1)
        mid = np.array([(edges[i - 1] + edges[i]) / 2 for i in range(1, len(edges))])
        return prob, mid

    def expr(self, F_samples, N, filt=0):
        """KL(P_pqc || P_haar(N)) of supplied fidelity samples (measure.py:161-180)."""
        if len(F_samples) == 0:
            return 0
        if isinstance(F_samples, torch.Tensor):
            F = F_samples.to(engine.device(), torch.float64)
        else:
            F = torch.as_tensor(np.asarray(F_samples, dtype=np.float64), device=engine.device())
        counts = engine.hist_f64(F, engine.n_bins(F.numel()))
        return float(engine.kl_haar(counts, N).item())

    def _expr_of_states(self, states, N):
        """Same value as expr(_gen_f_samples) without materialising the pair list."""
        S = states.shape[0]
        n_pairs = S * (S - 1) // 2
        if n_pairs == 0:
            return 0
        bins = engine.n_bins(n_pairs)
        if bins <= 0:
            raise ValueError("`bins` must be positive, when an integer")
        counts, _ = engine.fidelity_hist(states, bins=bins)
        return float(engine.kl_haar(counts, N).item())

    def expressibility(self, sample_N):
        N = 2 ** self.QC.n_qubits
        states = self._random_states(sample_N)
        if states is None:
            return 0
        return self._expr_of_states(states, N)

    def expressibility_streamed(self, sample_N, block, want_Q=False, resident_blocks=None,
                                checkpoint=None, stats=None):
        """expressibility(sample_N) -- and entanglement(sample_N) with want_Q -- for circuits
        whose states do not fit in memory together (additive API; BASELINE config 5).  Draws the
        same angle stream and hands blocks of `block` states to dist.streamed_expressibility:
        every rank keeps `resident_blocks` row blocks resident (default: what fits in 70 % of the
        free device memory, at most its share of all blocks), every column block is generated
        ONCE per round by its owner and broadcast to the other ranks over NVLink.
        `checkpoint` (path prefix) persists the int64 histogram and the position after every
        column block, so a multi-hour run resumes where it stopped with identical counts; the
        Meyer-Wallach values of a resumed run cover only the blocks generated after the restart
        (their owner keeps them in `<checkpoint>.rank<r>.Q.json`, merged on resume)."""
        from . import dist as pdist
        import json
        import os
        N = 2 ** self.QC.n_qubits
        if sample_N <= 0:
            return (0, []) if want_Q else 0
        ang = self.QC.draw_random(sample_N)
        rank, world = pdist.rank_world()
        if resident_blocks is None:
            nb = (sample_N + block - 1) // block
            share = (nb + world - 1) // world
            if torch.cuda.is_available():                          # free + what torch holds cached
                free = torch.cuda.mem_get_info()[0] + torch.cuda.memory_reserved() - \
                    torch.cuda.memory_allocated()
            else:
                free = 1 << 62
            fit = int(0.85 * free // (block * 16 * N)) - 4       # + two travelling, a prefetched and a fresh block
            resident_blocks = max(1, min(share, fit))
        Q = {}
        qfile = f"{checkpoint}.rank{rank}.Q.json" if checkpoint and want_Q else None
        if qfile and os.path.exists(qfile):
            Q = {int(k): v for k, v in json.load(open(qfile)).items()}

        def run_block(lo, hi):
            return self.QC.program.run(ang[lo:hi], init=self.QC.initial_state.tensor)

        def per_block(lo, hi, states):
            Q[lo] = engine.meyer_wallach(states).cpu().numpy().tolist()
            if qfile:
                with open(qfile + ".tmp", "w") as f:
                    json.dump(Q, f)
                os.replace(qfile + ".tmp", qfile)

        e = pdist.streamed_expressibility(run_block, sample_N, block, N,
                                          per_block=per_block if want_Q else None,
                                          resident_blocks=resident_blocks, checkpoint=checkpoint,
                                          stats=stats)
        if not want_Q:
            return e
        return e, [q for lo in sorted(Q) for q in Q[lo]]     # this rank's rows, in sample order

    def find_eff_H(self, circuit_f_samples, n):
        """Effective Hilbert-space dimension by minimising expr over N (measure.py:199-224)."""
        F = torch.as_tensor(np.asarray(circuit_f_samples, dtype=np.float64), device=engine.device())

        def wrapper(dim, F_samples=None):
            return self.expr(F, float(np.atleast_1d(dim)[0]), filt=0.2)

        def log_wrapper(dim, F_samples=None):
            return self.log_expr(F, dim)          # does not exist in the reference either (Q8)

        wrap_fn = log_wrapper if n > 10 else wrapper
        out = scipy.optimize.minimize(wrap_fn, [4], method="BFGS")
        if out.success is True:
            return out.x[0]
        print(out)
        return 0

    # ---- entanglement (measure.py:226-249) ---------------------------------------------------------
    def single_Q(self, system, n):
        st = _as_state(system)
        return float(engine.meyer_wallach(st.tensor.reshape(1, -1))[0].item())

    def entanglement(self, sample_N):
        states = self._random_states(sample_N)
        if states is None:
            return []
        return engine.meyer_wallach(states).cpu().numpy().tolist()

    # ---- magic (measure.py:251-368) --------------------------------------------------------------------
    def theta_to_magic(self, angles):
        return -1 * self.renyi_entropy_fast(self.QC.run(angles=angles))

    def theta_to_gkp(self, angles):
        return -1 * self.gkp_fast(self.QC.run(angles=angles))

    def numberToBase(self, n, b, n_qubits):
        digits = np.zeros(n_qubits, dtype=int)
        k = 0
        while n:
            digits[k] = int(n % b)
            n //= b
            k += 1
        return digits[::-1]

    def get_conversion_matrix_mod_add_index(self, base_states):
        idx = np.array([int("".join(str(int(d)) for d in b), 2) for b in base_states])
        return idx[:, None] ^ idx[None, :]

    def get_conversion_matrix_binary_prod(self, base_states):
        B = np.asarray(base_states)
        return (-1) ** np.mod(B @ B.T, 2)

    def get_conversion_matrices(self):
        """(xor table, sign table) of the reference's dense formulation
        (measure.py:304-313).  The FWHT kernel never needs them; they are provided for API
        compatibility and built only when asked (2 * 4^n * 8 bytes)."""
        n = self.QC.n_qubits
        base_states = [self.numberToBase(i, 2, n) for i in range(2 ** n)]
        return (self.get_conversion_matrix_mod_add_index(base_states),
                self.get_conversion_matrix_binary_prod(base_states))

    def set_converstion_matrices(self, conv_mats):
        self.conversion_matrices = conv_mats

    def renyi_entropy_fast(self, state, conversion_matrices=None, alpha=2):
        """Renyi-alpha stabilizer entropy; `conversion_matrices` is accepted and ignored."""
        st = _as_state(state)
        return float(engine.magic(st.tensor.reshape(1, -1), (alpha,))[0, 0].item())

    def entropy_of_magic(self, sample_N):
        states = self._random_states(sample_N)
        if states is None:                 # np.mean([]) in the reference (measure.py:356-359)
            return np.mean([])
        magics = engine.magic(states, (2.0,))[0].cpu().numpy()
        return np.mean(magics)

    def gkp_fast(self, state, conversion_matrices=None):
        return 1 / (2 * np.log(2)) * self.renyi_entropy_fast(state, conversion_matrices, alpha=1 / 2)

    # ---- everything at once (measure.py:370-459) --------------------------------------------------------
    def efficient_measurements(self, sample_N, measure_expr=True, measure_ent=True,
                               measure_eom=True, measure_GKP=True, full_data=False,
                               angles="random"):
        n = self.QC.n_qubits
        if sample_N == 0:
            measure_expr = measure_ent = measure_eom = measure_GKP = False
        if angles == "clifford":
            clifford_angles = (0, np.pi / 2, np.pi, 3 * np.pi / 2, 2 * np.pi)
            # the reference draws QC.n_params (= 2 x the true count, quirk Q1) per sample
            init_angles = [[random.choice(clifford_angles) for i in range(self.QC.n_params)]
                           for i in range(sample_N)]
            if sample_N > 0:
                P = self.QC.n_true_params
                states = self.QC.run_batch(np.asarray(init_angles, dtype=np.float64)[:, :P])
            else:
                states = None
        else:
            states = self._random_states(sample_N)

        overlaps, magics, gkps, q_vals = [], [], [], []
        if measure_expr and n < 12:
            S = states.shape[0]
            if full_data:
                _, F = engine.fidelity_hist(states, want_F=True)
                overlaps = F.cpu().numpy().tolist()
            if n < 7:
                expr = self._expr_of_states(states, 2 ** n) if S > 1 else 0
            else:
                expr = -1
        else:
            expr = -1

        if measure_ent:
            q_vals = engine.meyer_wallach(states).cpu().numpy().tolist()
            q, std = np.mean(q_vals), np.std(q_vals)
        else:
            q, std = -1, -1

        both = None
        if measure_eom or measure_GKP:
            both = engine.magic(states, (2.0, 0.5)).cpu().numpy()
        if measure_eom:
            magics = both[0].tolist()
            magic_bar, magic_std = np.mean(magics), np.std(magics)
        else:
            magic_bar, magic_std = -1, -1
        if measure_GKP:
            gkps = (both[1] / (2 * np.log(2))).tolist()
            gkp_bar, gkp_std = np.mean(gkps), np.std(gkps)
        else:
            gkp_bar, gkp_std = -1, -1

        if full_data is True:
            return {"Expr": overlaps, "Ent": q_vals, "Magic": magics, "GKP": gkps}
        return {"Expr": expr, "Ent": [q, std], "Magic": [magic_bar, magic_std],
                "GKP": [gkp_bar, gkp_std]}

    # ---- training (measure.py:461-553) -----------------------------------------------
    def get_gradient_vector(self, theta):
        """d_i <H> = 2 Re <psi|H|d_i psi> for every parameter."""
        self.QC.state = self.QC.run(angles=theta)
        psi = self.QC.state
        self.gradient_list = self.QC.get_gradients()
        if not self.gradient_list:
            return []
        grads = _stack(self.gradient_list)
        Hd = engine.pauli_apply(grads, self.QC.H.device_terms())
        ov = engine.overlap(psi.tensor.reshape(1, -1), Hd)
        return (2 * ov.real).cpu().numpy().tolist()

    def train(self, epsilon=1e-6, rate=0.001, method="gradient", angles="random", verbose=False):
        """Minimise the current objective; returns (energy, traj, magics, ents, gkps)."""
        quit_iterations = 100000
        count, diff = 0, 1
        traj, magics, gkps, ents = [], [], [], []

        def trajmaj(Xi):
            magics.append(self.renyi_entropy_fast(self.QC.state))
            traj.append(self.minimize_function(Xi))
            ents.append(self.single_Q(self.QC.state, self.QC.n_qubits))
            gkps.append(self.gkp_fast(self.QC.state))

        self.QC.state = self.QC.run(angles=angles)
        trajmaj(angles)

        if method.lower() in ["gradient", "qng"]:
            prev_energy = self.minimize_function(angles)
            while diff > epsilon and count < quit_iterations:
                theta = self.QC.get_params_raw() if hasattr(self.QC, "get_params_raw") \
                    else self.QC.get_params()
                gradients = self.get_gradient_vector(theta)
                if method == "gradient":
                    theta_update = list(np.array(theta) - rate * np.array(gradients))
                elif method == "QNG":
                    QFI = self.get_QFI(grad_list=self.gradient_list)
                    theta_update = list(np.array(theta) -
                                        rate * np.linalg.pinv(QFI).dot(np.array(gradients)))
                if count % 100 == 0 and verbose is True:
                    print(f"On iteration {count}, energy = {prev_energy}, diff is {diff}")
                energy = self.minimize_function(theta_update)
                diff = np.abs(energy - prev_energy)
                trajmaj(theta_update)
                count += 1
                prev_energy = energy
        else:
            op_out = scipy.optimize.minimize(self.minimize_function, x0=angles, method=method,
                                             callback=trajmaj, tol=epsilon)
            energy = op_out.fun
        return (energy, traj, magics, ents, gkps)

    # ---- additive batch entry points ---------------------------------------------------------------------
    def qfim_batch(self, angles, cutoff_eigvals=None, want_eigvals=False):
        """QFIM [S,P,P] (device) for every row of angles; with a cutoff also the effective
        quantum dimensions [S] (int32, device), and with `want_eigvals` the ascending spectra
        [S,P] they were counted from (get_eigenvalues, measure.py:73-75)."""
        F = self.QC.qfim_batch(angles)
        if cutoff_eigvals is None and not want_eigvals:
            return F
        w = engine.eigvalsh(F)
        out = (F,)
        if cutoff_eigvals is not None:
            out += (engine.count_greater(w, cutoff_eigvals),)
        if want_eigvals:
            out += (w,)
        return out
