"""CPU ORACLE for the SCoNe hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (scone_gcn_b200) never does; it fails loudly without its CUDA library.

Parity status: the reference (nglaze00/SCoNe_GCN) ships NO tests or golden vectors, so
parity is pinned the other way the task allows: tests/golden/*.npz are outputs of the reference's own
Python executed unmodified in the build container by oracle/make_golden.py (data/integer code on the
real NumPy/SciPy/NetworkX stack; model code over a torch-backed `jax` stand-in because jax is not
installable offline).  tests/test_oracle_golden.py checks every function below against those files.

Everything here restates, op for op, these reference locations (paths relative to
/root/reference/trajectory_analysis/):
  incidence_matrices        synthetic_data_gen.py:139-161
  path_to_flow              synthetic_data_gen.py:327-344
  neighborhood_to_onehot    synthetic_data_gen.py:288-297
  shift assembly            trajectory_experiments.py:239-260
  nbrhoods / n_nbrs / B1_jax / Bconds_func   trajectory_experiments.py:262-303
  bunch operators           bunch_model_matrices.py:44-135
  scone_func / ebli_func / bunch_func        trajectory_experiments.py:137-203
  generate_weights          scone_trajectory_model.py:215-242 (+ seed at :15)
  loss / accuracy           scone_trajectory_model.py:42-71
  two_target_accuracy       scone_trajectory_model.py:73-108
  train / Adam              scone_trajectory_model.py:264-357 (+ upstream JAX adam formula)
Two arithmetic back-ends:
  DenseOracle   torch CPU, dense E x E operators, `(S @ H) @ W` association, autograd — the reference
                formulation; fp32 (JAX default) or fp64 (tolerance budgeting).
  SparseOracle  NumPy + SciPy CSR with a hand-derived backward — the same maths
                at sizes where a dense E x E operator cannot exist (configs 4-5); validated against
                DenseOracle in tests/test_oracle_golden.py.  scone / ebli only.
"""
import numpy as np
import scipy.sparse as sp
import torch

# ------------------------------------------------------------------------------------------------
# integer / index work (bit-exact contract)
# ------------------------------------------------------------------------------------------------


def incidence_matrices(n_nodes, edges, faces):
    """Dense B1 [N,E], B2 [E,F] (float64, integer-valued).  synthetic_data_gen.py:139-161.

    B1[tail,e] = -1, B1[head,e] = +1 with tail<head (nx.incidence_matrix(oriented=True) on edges (a,b),
    a<b).  B2[(a,b),f] = B2[(b,c),f] = +1, B2[(a,c),f] = -1 for a sorted face (a<b<c)."""
    edges = np.asarray(edges).reshape(-1, 2)
    faces = np.asarray(faces).reshape(-1, 3)
    E, F = len(edges), len(faces)
    B1 = np.zeros((n_nodes, E))
    B1[edges[:, 0], np.arange(E)] = -1.0
    B1[edges[:, 1], np.arange(E)] = 1.0
    lut = {(int(a), int(b)): i for i, (a, b) in enumerate(edges)}
    B2 = np.zeros((E, F))
    for j, (a, b, c) in enumerate(faces):
        B2[lut[(int(a), int(b))], j] = 1.0
        B2[lut[(int(b), int(c))], j] = 1.0
        B2[lut[(int(a), int(c))], j] = -1.0
    return B1, B2


def path_to_flow(path, edge_to_idx, m):
    """synthetic_data_gen.py:327-344: +1 when traversed low->high node id, -1 otherwise."""
    f = np.zeros((m, 1))
    for v0, v1 in zip(path[:-1], path[1:]):
        if v0 < v1:
            f[edge_to_idx[(int(v0), int(v1))]] += 1
        else:
            f[edge_to_idx[(int(v1), int(v0))]] -= 1
    return f


def neighborhood_to_onehot(Nv, w, D):
    """synthetic_data_gen.py:288-297."""
    onehot = (np.asarray(Nv) == w).astype(float)
    out = np.zeros(D)
    out[:onehot.shape[0]] = onehot
    return np.array([out]).T


def adjacency_from_B1(B1):
    """Sorted neighbour lists from the nonzero pattern of B1 (as E_lookup is derived at
    trajectory_experiments.py:263-268)."""
    N, E = B1.shape
    nbrs = [[] for _ in range(N)]
    cols = np.nonzero(B1.T)
    ends = cols[1].reshape(-1, 2)
    for a, b in ends:
        nbrs[int(a)].append(int(b))
        nbrs[int(b)].append(int(a))
    return [sorted(set(v)) for v in nbrs]


def neighbourhood_tables(B1, last_nodes):
    """nbrhoods [N,D] (sorted ascending, pad -1), n_nbrs per trajectory, B1_jax (B1 + zero row).
    trajectory_experiments.py:272-288."""
    nbrs = adjacency_from_B1(B1)
    D = max(len(v) for v in nbrs)
    nbrhoods = np.array([v + [-1] * (D - len(v)) for v in nbrs], dtype=np.int64)
    n_nbrs = np.array([len(nbrs[int(n)]) for n in last_nodes], dtype=np.int64)
    B1_jax = np.append(B1, np.zeros((1, B1.shape[1])), axis=0)
    return nbrhoods, n_nbrs, B1_jax


def shift_matrices(B1, B2, model, flips=None):
    """trajectory_experiments.py:239-260 (flips: :214-219,242-244)."""
    L_lower = B1.T @ B1
    L_upper = B2 @ B2.T
    if flips is not None:
        Fm = np.diag(flips)
        L_lower = Fm @ L_lower @ Fm
        L_upper = Fm @ L_upper @ Fm
    if model == 'scone':
        return [L_lower, L_upper]
    if model == 'ebli':
        L1 = L_lower + L_upper
        return [L1, L1 @ L1]
    if model == 'bunch':
        return list(bunch_shift_matrices(B1, B2))
    raise Exception('invalid model type')


def bunch_shift_matrices(B1, B2):
    """bunch_model_matrices.py:44-135, same dense float64 operations in the same order."""
    from numpy.linalg import inv, pinv
    D2_2 = np.diag(np.maximum(np.abs(B2).sum(axis=1), 1))          # compute_D2(B2)  :44-52
    D2_1 = np.diag(np.maximum(np.abs(B1).sum(axis=1), 1))          # compute_D2(B1)
    D3_n = np.identity(B1.shape[1])
    D1 = 2 * np.diag((np.abs(B1) @ D2_2).sum(axis=1))              # compute_D1      :62-69
    D3 = np.identity(B2.shape[1]) / 3
    D4 = np.identity(B2.shape[1])
    D5 = np.diag(np.abs(B2).sum(axis=1))                           # compute_D5      :53-60
    D1_pinv, D5_pinv, D2_2_inv = pinv(D1), pinv(D5), inv(D2_2)
    L0u = B1 @ D3_n @ B1.T @ inv(D2_1)
    L1u = D2_2 @ B1.T @ D1_pinv @ B1
    L1d = B2 @ D3 @ B2.T @ D2_2_inv
    L2d = D4 @ B2.T @ D5_pinv @ B2
    D4_inv = inv(D4)
    A0u = D2_1 - (L0u @ D2_1)
    A1u = D2_2 - (L1u @ D2_2)
    A1d = D2_2_inv - (D2_2_inv @ L1d)
    A2d = D4_inv - (D4_inv @ L2d)
    I0, I1, I2 = np.identity(A0u.shape[0]), np.identity(A1u.shape[0]), np.identity(A2d.shape[0])
    A0u_n = (A0u + I0) @ inv(D2_1 + I0)
    A1u_n = (A1u + I1) @ inv(D2_2 + I1)
    A1d_n = (D2_2 + I1) @ (A1d + I1)
    A2d_n = (D4 + I2) @ (A2d + I2)
    return (A0u_n, D1_pinv @ B1, D2_2 @ B1.T @ D1_pinv, A1d_n + A1u_n, B2 @ D3,
            D4 @ B2.T @ D5_pinv, A2d_n)


def weight_shapes(in_channels, hidden_layers, out_channels, model_type):
    """scone_trajectory_model.py:222-234."""
    shapes = [(in_channels, hidden_layers[0][1])] * hidden_layers[0][0]
    for i in range(len(hidden_layers) - 1):
        shapes += [(hidden_layers[i][1], hidden_layers[i + 1][1])] * hidden_layers[i + 1][0]
    if model_type == 'bunch':
        shapes += [(hidden_layers[-1][1], out_channels)] * hidden_layers[-1][0]
    else:
        shapes += [(hidden_layers[-1][1], out_channels)]
    return shapes


def generate_weights(rng, in_channels, hidden_layers, out_channels, model_type):
    """0.01 * randn per weight, in list order, from the given legacy RandomState stream
    (scone_trajectory_model.py:15,236-237)."""
    return [0.01 * rng.randn(*s) for s in weight_shapes(in_channels, hidden_layers, out_channels, model_type)]


# ------------------------------------------------------------------------------------------------
# dense reference formulation (torch CPU, autograd)
# ------------------------------------------------------------------------------------------------
_ACT = {
    'scone': torch.tanh,                                                   # trajectory_experiments.py:130-131
    'ebli': lambda x: torch.where(x >= 0, x, 0.01 * x),                    # :133-134
    'bunch': lambda x: torch.maximum(x, torch.zeros((), dtype=x.dtype)),   # :124-125
}


class DenseOracle:
    """The reference formulation, batched: dense shifts, `(S @ H) @ W`, padded-zero log-softmax."""

    def __init__(self, model, shifts, B1, last_nodes, flows, targets, dtype=torch.float32):
        self.model = model
        self.dt = dtype
        self.shifts = [torch.as_tensor(np.asarray(s), dtype=dtype) for s in shifts]
        nbrhoods, n_nbrs, B1_jax = neighbourhood_tables(np.asarray(B1), last_nodes)
        self.nbrhoods = torch.as_tensor(nbrhoods)
        self.n_nbrs = n_nbrs
        self.B1_jax = torch.as_tensor(B1_jax, dtype=dtype)
        self.last_nodes = torch.as_tensor(np.asarray(last_nodes), dtype=torch.int64)
        self.flows = torch.as_tensor(np.asarray(flows), dtype=dtype).reshape(len(last_nodes), -1, 1)
        self.y = torch.as_tensor(np.asarray(targets), dtype=dtype).reshape(len(last_nodes), -1, 1)
        self.n_shifts = len(self.shifts)

    def _w(self, weights):
        return [w if isinstance(w, torch.Tensor) else torch.as_tensor(np.asarray(w), dtype=self.dt) for w in weights]

    def forward(self, weights, idx=None):
        """log-probs [B, D, 1].  trajectory_experiments.py:137-203."""
        W = self._w(weights)
        act = _ACT[self.model]
        X = self.flows if idx is None else self.flows[idx]
        last = self.last_nodes if idx is None else self.last_nodes[idx]
        Nv = self.nbrhoods[last]                                             # [B, D]
        if self.model in ('scone', 'ebli'):
            assert (len(W) - 1) % 3 == 0, 'wrong number of weights'
            S0, S1 = self.shifts
            cur = X
            for i in range((len(W) - 1) // 3):
                cur = cur @ W[3 * i] + (S0 @ cur) @ W[3 * i + 1] + (S1 @ cur) @ W[3 * i + 2]
                cur = act(cur)
            Bcond = self.B1_jax[Nv]                                          # [B, D, E]; -1 -> zero row
            logits = (Bcond @ cur) @ W[-1]
        else:
            assert len(W) % 7 == 0, 'wrong number of weights'
            S_00, S_10, S_01, S_11, S_21, S_12, S_22 = self.shifts
            B = X.shape[0]
            cur = [torch.zeros(B, S_00.shape[1], 1, dtype=self.dt), X, torch.zeros(B, S_22.shape[1], 1, dtype=self.dt)]
            for i in range(len(W) // 7):
                n0 = (S_00 @ cur[0]) @ W[7 * i] + (S_10 @ cur[1]) @ W[7 * i + 1]
                n1 = (S_01 @ cur[0]) @ W[7 * i + 2] + (S_11 @ cur[1]) @ W[7 * i + 3] + (S_21 @ cur[2]) @ W[7 * i + 4]
                n2 = (S_12 @ cur[1]) @ W[7 * i + 5] + (S_22 @ cur[2]) @ W[7 * i + 6]
                cur = [act(n0), act(n1), act(n2)]
            nodes_out = cur[0]                                               # [B, N, 1]
            logits = torch.gather(nodes_out, 1, (Nv % nodes_out.shape[1]).unsqueeze(-1))   # -1 wraps to N-1 (Q2)
        return logits - torch.logsumexp(logits, dim=1, keepdim=True)

    def loss(self, weights, mask, wd):
        """scone_trajectory_model.py:42-56 (forward over ALL trajectories, then mask-select)."""
        W = self._w(weights)
        m = torch.as_tensor(np.asarray(mask) == 1)
        preds = self.forward(W)[m]
        k = self.n_shifts + 1 if self.model != 'bunch' else self.n_shifts
        last = W[-1:] if self.model != 'bunch' else W[-k:]
        mid = W[k:-1] if self.model != 'bunch' else W[k:-k]
        ridge = sum((w ** 2).sum() for w in W[:k]) + sum((w ** 2).sum() for w in mid) + sum((w ** 2).sum() for w in last)
        return -(preds * self.y[m]).sum() / m.sum() + wd * ridge

    def loss_and_grads(self, weights, mask, wd):
        W = [torch.as_tensor(np.asarray(w), dtype=self.dt).requires_grad_(True) for w in weights]
        l = self.loss(W, mask, wd)
        g = torch.autograd.grad(l, W)
        return l.detach().numpy(), [x.numpy() for x in g]

    def accuracy(self, weights, mask):
        """scone_trajectory_model.py:59-71."""
        with torch.no_grad():
            preds = self.forward(weights).numpy().copy()
        for i in range(len(preds)):
            preds[i, self.n_nbrs[i]:] = -100
        m = np.asarray(mask) == 1
        return float(np.mean(np.argmax(preds[m], axis=1) == np.argmax(self.y.numpy()[m], axis=1)))


def two_target_accuracy(preds, y, mask, n_nbrs, random_targets=None, rng=np.random):
    """scone_trajectory_model.py:73-108 on the model's log-probs `preds` [N, D, 1]: returns (score, random_targets).

    `rng` is the legacy global NumPy stream the reference draws from (:79, :91).  pred_choice is a jax array there, so
    pred_choice[i] with i past the end clamps to the last element: the redraw loop runs over all N rows (:89-91)."""
    preds = np.array(preds, dtype=np.float32).reshape(len(preds), -1)
    mask = np.asarray(mask)
    n_nbrs = np.asarray(n_nbrs)
    N = len(preds)
    if random_targets is None:
        random_targets = rng.randint(0, high=n_nbrs, size=N)
    for i in range(N):
        preds[i, n_nbrs[i]:] = -100
    pred_choice = np.argmax(preds[mask == 1], axis=1)
    last = len(pred_choice) - 1
    for i in range(N):
        while random_targets[i] == pred_choice[min(i, last)]:
            random_targets[i] = rng.randint(0, high=n_nbrs[i])
    rows = np.arange(N)
    random_probs = preds[rows, random_targets]
    true_probs = preds[rows, np.argmax(np.asarray(y).reshape(N, -1), axis=1)]
    correct = 0.0
    for t, r in zip(true_probs[mask == 1], random_probs[mask == 1]):
        if t > r:
            correct += 1
        elif t == r:
            correct += 0.5
    return correct / mask.sum(), random_targets


def adam_update(i, grads, state, step_size, b1=0.9, b2=0.999, eps=1e-8, dtype=np.float32):
    """Upstream JAX `adam` (not under /root/reference; used at scone_trajectory_model.py:300,310,325-326)."""
    new = []
    f = dtype
    for g, (x, m, v) in zip(grads, state):
        g = np.asarray(g, dtype=f)
        m = f(1 - b1) * g + f(b1) * m
        v = f(1 - b2) * g * g + f(b2) * v
        mhat = m / (f(1) - f(b1) ** f(i + 1))
        vhat = v / (f(1) - f(b2) ** f(i + 1))
        x = x - f(step_size) * mhat / (np.sqrt(vhat) + f(eps))
        new.append((x.astype(f), m.astype(f), v.astype(f)))
    return new


def train(oracle, rng, weights, train_mask, test_mask, epochs, batch_size, step_size, wd, dtype=np.float32):
    """scone_trajectory_model.py:264-357: steps = epochs*(n_train//bs); batch mask = shuffle ∧ train."""
    N = oracle.flows.shape[0]
    n_batches = int(sum(train_mask)) // batch_size
    state = [(np.asarray(w, dtype=dtype), np.zeros_like(w, dtype=dtype), np.zeros_like(w, dtype=dtype)) for w in weights]
    unshuffled = np.array([1] * batch_size + [0] * (N - batch_size))
    W = [s[0] for s in state]
    for i in range(epochs * n_batches):
        bm = np.array(unshuffled)
        rng.shuffle(bm)
        bm = np.logical_and(bm, train_mask)
        _, g = oracle.loss_and_grads(W, bm, wd)
        state = adam_update(i, g, state, step_size, dtype=dtype)
        W = [s[0] for s in state]
    with torch.no_grad():
        res = (float(oracle.loss(W, train_mask, wd)), oracle.accuracy(W, train_mask),
               float(oracle.loss(W, test_mask, wd)), oracle.accuracy(W, test_mask))
    return W, res


# ------------------------------------------------------------------------------------------------
# sparse formulation with hand-derived backward (scales to configs 4-5; scone / ebli)
# ------------------------------------------------------------------------------------------------
class SparseOracle:
    """Same maths with CSR operators (torch CPU sparse, multi-threaded), layout H[E, b, C] (edge-major),
    hand-derived backward.  Used (a) to cross-check the CUDA kernels at sizes where a
    dense E x E operator cannot exist and (b) as the CPU baseline ('port') timed by bench.py."""

    def __init__(self, model, edges, tri_edges, tri_signs, n_nodes, dtype=np.float32):
        """edges [E,2] (tail<head); tri_edges [F,3] edge ids; tri_signs [F,3] in {+1,-1}."""
        self.model, self.dt = model, dtype
        self.tdt = torch.float32 if dtype == np.float32 else torch.float64
        edges = np.asarray(edges)
        E, F = len(edges), len(tri_edges)
        self.E, self.N = E, n_nodes
        ar = np.arange(E)
        B1 = sp.csr_matrix((np.r_[-np.ones(E), np.ones(E)], (np.r_[edges[:, 0], edges[:, 1]], np.r_[ar, ar])),
                           shape=(n_nodes, E))
        B2 = sp.csr_matrix((np.asarray(tri_signs, dtype=np.float64).ravel(),
                            (np.asarray(tri_edges).ravel(), np.repeat(np.arange(F), 3))), shape=(E, F))
        Ll, Lu = (B1.T @ B1).tocsr(), (B2 @ B2.T).tocsr()
        if model == 'scone':
            S = [Ll, Lu]
        else:
            L1 = (Ll + Lu).tocsr()
            L1.eliminate_zeros()
            S = [L1, (L1 @ L1).tocsr()]
        self.S = [self._tcsr(s) for s in S]
        self.B1 = self._tcsr(B1.tocsr())
        self.B1T = self._tcsr(B1.T.tocsr())
        adj = (abs(B1) @ abs(B1).T).tolil()
        adj.setdiag(0)
        adj = adj.tocsr()
        adj.eliminate_zeros()
        adj.sort_indices()
        self.nbr_ptr, self.nbr_idx = adj.indptr, adj.indices           # sorted ascending per row
        self.D = int(np.diff(adj.indptr).max())

    def _tcsr(self, m):
        m.sort_indices()
        return torch.sparse_csr_tensor(torch.as_tensor(m.indptr.astype(np.int64)), torch.as_tensor(m.indices.astype(np.int64)),
                                       torch.as_tensor(m.data.astype(self.dt)), size=m.shape)

    def act(self, z):
        if self.model == 'scone':
            return torch.tanh(z)
        return torch.where(z >= 0, z, 0.01 * z)

    def dact(self, h):
        if self.model == 'scone':
            return 1 - h * h
        one = torch.ones((), dtype=h.dtype)
        return torch.where(h >= 0, one, 0.01 * one)

    def _shift(self, k, H):
        E, b, C = H.shape
        return (self.S[k] @ H.reshape(E, b * C)).reshape(E, b, C)

    def forward(self, weights, X, last_nodes, keep=False):
        """X [E, b] dense flows (edge-major).  Returns log-probs [b, D] (+ saved activations)."""
        with torch.no_grad():
            W = [torch.as_tensor(np.asarray(w, dtype=self.dt)) for w in weights]
            X = torch.as_tensor(np.asarray(X, dtype=self.dt))
            E, b = X.shape
            L = (len(W) - 1) // 3
            H = X.reshape(E, b, 1)
            acts = [H]
            for i in range(L):
                H = self.act(H @ W[3 * i] + self._shift(0, H) @ W[3 * i + 1] + self._shift(1, H) @ W[3 * i + 2])
                acts.append(H)
            q = (H @ W[-1])[:, :, 0]                                        # [E, b]
            div = (self.B1 @ q).numpy()                                     # [N, b]
            logits = np.zeros((b, self.D), dtype=self.dt)
            for t, n in enumerate(last_nodes):
                nb = self.nbr_idx[self.nbr_ptr[n]:self.nbr_ptr[n + 1]]
                logits[t, :len(nb)] = div[nb, t]
            mx = logits.max(axis=1, keepdims=True)
            lse = mx + np.log(np.exp(logits - mx).sum(axis=1, keepdims=True))
            lp = logits - lse
        return (lp, acts, W) if keep else lp

    def loss_and_grads(self, weights, X, last_nodes, target_idx, mask, n_total=None):
        """Returns (nll_sum, [dW]) WITHOUT ridge; gradients are scaled by 1/n_total when it is given."""
        lp, acts, W = self.forward(weights, X, last_nodes, keep=True)
        with torch.no_grad():
            E, b = acts[0].shape[:2]
            L = (len(W) - 1) // 3
            mask = np.asarray(mask, dtype=self.dt)
            scale = self.dt(1.0 if n_total is None else 1.0 / n_total)
            nll = -float((lp[np.arange(b), target_idx] * mask).sum())
            dlogit = np.exp(lp)
            dlogit[np.arange(b), target_idx] -= 1
            dlogit *= (mask * scale)[:, None]
            ddiv = np.zeros((self.N, b), dtype=self.dt)
            for t, n in enumerate(last_nodes):
                nb = self.nbr_idx[self.nbr_ptr[n]:self.nbr_ptr[n + 1]]
                ddiv[nb, t] = dlogit[t, :len(nb)]
            dq = self.B1T @ torch.as_tensor(ddiv)                           # [E, b]
            HL = acts[-1]
            grads = [None] * len(W)
            grads[-1] = torch.einsum('ebc,eb->c', HL, dq).reshape(-1, 1).numpy()
            dH = dq[:, :, None] * W[-1][:, 0][None, None, :]
            for i in range(L - 1, -1, -1):
                Hin, Hout = acts[i], acts[i + 1]
                G = dH * self.dact(Hout)
                Cin, Cout = Hin.shape[2], Hout.shape[2]
                A1, A2 = self._shift(0, G), self._shift(1, G)
                Hf = Hin.reshape(E * b, Cin)
                grads[3 * i] = (Hf.T @ G.reshape(E * b, Cout)).numpy()
                grads[3 * i + 1] = (Hf.T @ A1.reshape(E * b, Cout)).numpy()
                grads[3 * i + 2] = (Hf.T @ A2.reshape(E * b, Cout)).numpy()
                if i > 0:
                    dH = G @ W[3 * i].T + A1 @ W[3 * i + 1].T + A2 @ W[3 * i + 2].T
        return nll, grads
