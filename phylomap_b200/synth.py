"""Synthetic inputs of the shapes BASELINE.json names: trees, rate matrices and tip data.

Generalises the reference's simulators, which hard-code the tip count (R/simulate_2_state_tree.R:12,17,29 uses
3951; R/sourceme.R:419-432 uses 70): `simulate_2_state_tree` / `simulate_4_state_tree` here work for any tree and
any number of sites.  Tip data are drawn with the exact transition probabilities P(t) = expm(Q t) level by level
(same law as the Gillespie walk of sample2statehistory, R/sourceme.R:375-410, without storing the paths).
"""
import numpy as np

from .tree import PhyloTree


def make2sQ(l01, l10, rkappas, lkappas, gammas):
    """Hidden-rate 2(k+1)-state generator, R/sourceme.R:229-246 (states 1-based there, 0-based rows here)."""
    rk, lk, ga = np.atleast_1d(rkappas), np.atleast_1d(lkappas), np.atleast_1d(gammas)
    k = len(rk)
    n = 2 * k + 2
    Q = np.zeros((n, n))
    Q[0, 1], Q[1, 0] = l01, l10
    for i in range(1, k + 1):
        Q[2 * i - 2, 2 * i] = rk[i - 1]
        Q[2 * i - 1, 2 * i + 1] = rk[i - 1]
        Q[2 * i, 2 * i - 2] = lk[i - 1]
        Q[2 * i + 1, 2 * i - 1] = lk[i - 1]
        Q[2 * i, 2 * i + 1] = ga[i - 1] * l01
        Q[2 * i + 1, 2 * i] = ga[i - 1] * l10
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    return Q


def _relabel_cladewise(parent_of, children, root, T, blen):
    """ape numbering: tips 1..T, root T+1, internal nodes in preorder; edges in preorder (cladewise)."""
    new_id = {}
    next_tip, next_int = 1, T + 1
    edges, lens = [], []
    stack = [root]
    order = []
    while stack:
        v = stack.pop()
        order.append(v)
        if children[v]:
            new_id[v] = next_int
            next_int += 1
            stack.append(children[v][1])
            stack.append(children[v][0])
        else:
            new_id[v] = next_tip
            next_tip += 1
    for v in order:
        if v != root:
            edges.append((new_id[parent_of[v]], new_id[v]))
            lens.append(blen[v])
    return np.array(edges, dtype=np.int32), np.array(lens, dtype=np.float64)


def yule_tree(T, seed=1, birth=0.1, mean_branch=None):
    """Pure-birth tree with T tips.  mean_branch rescales the branch lengths to that mean."""
    rng = np.random.default_rng(seed)
    parent_of, children, born, blen = {0: -1}, {0: []}, {0: 0.0}, {}
    active = [0]
    t = 0.0
    nxt = 1
    while len(active) < T:
        t += rng.exponential(1.0 / (birth * len(active)))
        i = int(rng.integers(len(active)))
        v = active[i]
        blen[v] = t - born[v]
        a, b = nxt, nxt + 1
        nxt += 2
        for c in (a, b):
            parent_of[c] = v
            children[c] = []
            born[c] = t
        children[v] = [a, b]
        active[i] = a
        active.append(b)
    t += rng.exponential(1.0 / (birth * len(active)))
    for v in active:
        blen[v] = t - born[v]
    edge, el = _relabel_cladewise(parent_of, children, 0, T, blen)
    if mean_branch is not None:
        el = el * (mean_branch / el.mean())
    return PhyloTree(edge, el)


def balanced_tree(depth, branch=1.0):
    T = 1 << depth
    parent_of, children, blen = {0: -1}, {0: []}, {}
    nxt = 1
    frontier = [0]
    for _ in range(depth):
        new = []
        for v in frontier:
            a, b = nxt, nxt + 1
            nxt += 2
            children[v] = [a, b]
            for c in (a, b):
                parent_of[c] = v
                children[c] = []
                blen[c] = branch
                new.append(c)
        frontier = new
    edge, el = _relabel_cladewise(parent_of, children, 0, T, blen)
    return PhyloTree(edge, el)


def read_newick(text):
    """Minimal newick reader for strictly binary trees with branch lengths (e.g. inst/extdata/Squamate/squamate.phy)."""
    s = text.strip()
    if s.endswith(";"):
        s = s[:-1]
    parent_of, children, blen = {}, {}, {}

    def new_node(par):
        v = len(parent_of)
        parent_of[v] = par
        children[v] = []
        if par >= 0:
            children[par].append(v)
        return v

    label = {}

    def set_len(v, tok):
        if ":" in tok:
            name, length = tok.rsplit(":", 1)
            blen[v] = float(length)
        else:
            name = tok
        label[v] = name.strip().strip("'")

    stack, i, root = [], 0, None
    while i < len(s):
        c = s[i]
        if c == "(":
            v = new_node(stack[-1] if stack else -1)
            if root is None:
                root = v
            stack.append(v)
            i += 1
        elif c == ",":
            i += 1
        else:
            closing = c == ")"
            if closing:
                i += 1
            j = i
            while j < len(s) and s[j] not in ",()":
                j += 1
            v = stack.pop() if closing else new_node(stack[-1])
            set_len(v, s[i:j].strip())
            i = j
    for v, ch in children.items():
        if len(ch) not in (0, 2):
            raise ValueError("newick tree is not binary")
    T = sum(1 for v in children if not children[v])
    for v in parent_of:
        blen.setdefault(v, 0.0)
    edge, el = _relabel_cladewise(parent_of, children, root, T, blen)
    # tips are numbered in order of appearance (ape::read.tree): the same walk _relabel_cladewise does
    tips, stack2 = [], [root]
    while stack2:
        v = stack2.pop()
        if children[v]:
            stack2.append(children[v][1])
            stack2.append(children[v][0])
        else:
            tips.append(label.get(v, ""))
    return PhyloTree(edge, el, tip_label=tips)


def make_squamate_tree(newick_path, tipdata_csv, segments=100):
    """R/Squamate_tree_setup.R:8-88 (`make_squamate_tree`): the Squamate phylogeny in the form the samplers take -- every
    branch cut into 100 equal segments named 1, tip states from the trait table ("0" -> 1, "1" -> 2, anything else keeps the
    script's placeholder -10), the last segment of a tip branch named after its tip state.  The tree file shipped in
    inst/extdata already holds the 3 951 tips the script keeps."""
    import csv
    with open(newick_path) as f:
        tree = read_newick(f.read())
    with open(tipdata_csv, newline="") as f:
        rows = list(csv.reader(f))[1:]
    parity = {r[0]: r[1] for r in rows}
    code = {"0": 1, "1": 2}
    states = np.array([code.get(parity.get(name, ""), -10) for name in tree.tip_label], dtype=np.int32)
    maps = [np.full(segments, t / segments) for t in tree.edge_length]
    names = [np.ones(segments, dtype=np.int32) for _ in range(tree.E)]
    for e in range(tree.E):
        c = tree.edge[e, 1]
        if c <= tree.T:
            names[e][-1] = states[c - 1]
    return PhyloTree(tree.edge, tree.edge_length, states, maps, names, tree.tip_label)


def _expm(Qt):
    import torch
    return torch.matrix_exp(Qt)


def simulate_tip_states(tree, Q, pid, n_sites, seed=101, device="cpu", batch_sites=None):
    """Tip states [S, T] uint8, 1-based, of S independent characters evolved down `tree` under Q from a root
    drawn from pid.  Runs on `device` through torch (plumbing only: synthetic data, not the sampler)."""
    import torch
    T, E = tree.T, tree.E
    n = Q.shape[0]
    parent = tree.edge[:, 0] - 1
    child = tree.edge[:, 1] - 1
    NN = 2 * T - 1
    root = int(np.setdiff1d(parent, child)[0])
    # depth levels of edges
    pe = np.full(NN, -1, dtype=np.int64)
    pe[child] = np.arange(E)
    depth = np.zeros(NN, dtype=np.int64)
    order = np.argsort(parent, kind="stable")
    kids = {}
    for e in range(E):
        kids.setdefault(parent[e], []).append(e)
    levels = []
    frontier = [root]
    while frontier:
        es = [e for v in frontier for e in kids.get(v, [])]
        if not es:
            break
        levels.append(np.array(es, dtype=np.int64))
        frontier = [child[e] for e in es]
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    Qt = torch.as_tensor(Q, dtype=torch.float64, device=dev)
    el = torch.as_tensor(tree.edge_length, dtype=torch.float64, device=dev)
    pid_t = torch.as_tensor(np.asarray(pid, dtype=np.float64), device=dev)
    out = torch.empty((n_sites, T), dtype=torch.uint8, device=dev)
    bs = batch_sites or n_sites
    for s0 in range(0, n_sites, bs):
        ns = min(bs, n_sites - s0)
        st = torch.zeros((NN, ns), dtype=torch.int64, device=dev)
        u = torch.rand(ns, generator=g, device=dev, dtype=torch.float64)
        st[root] = torch.searchsorted(torch.cumsum(pid_t / pid_t.sum(), 0), u).clamp_(max=n - 1)
        for es in levels:
            es_t = torch.as_tensor(es, device=dev)
            P = _expm(Qt[None, :, :] * el[es_t][:, None, None]).clamp_(min=0)  # [k, n, n]
            cdf = torch.cumsum(P, dim=2)
            cdf = cdf / cdf[:, :, -1:]
            ps = st[torch.as_tensor(parent[es], device=dev)]  # [k, ns]
            rows = torch.gather(cdf, 1, ps[:, :, None].expand(-1, -1, n))  # [k, ns, n]
            u = torch.rand(rows.shape[:2], generator=g, device=dev, dtype=torch.float64)
            cs = (u[:, :, None] > rows).sum(dim=2).clamp_(max=n - 1)
            st[torch.as_tensor(child[es], device=dev)] = cs
        out[s0:s0 + ns] = (st[:T].t() + 1).to(torch.uint8)
    return out


def simulate_2_state_tree(seed, atree, Q, pid2, n_sites=1, device="cpu", segments=None):
    """R/simulate_2_state_tree.R:8-32 for any tree / any number of sites: simulate tips, halve the tip branches."""
    st = simulate_tip_states(atree, np.asarray(Q, dtype=np.float64), pid2, n_sites, seed, device).cpu().numpy()
    return atree.with_states(st[0].astype(np.int32) if n_sites == 1 else st, segments=segments)


def simulate_4_state_tree(seed, atree, Q, pid4, n_sites=1, device="cpu", segments=None):
    """R/simulate_4_state_tree.R:7-34: simulate under the hidden-rate Q, observe the trait (odd states -> 1,
    even states -> 2 in the reference's 1-based numbering)."""
    st = simulate_tip_states(atree, np.asarray(Q, dtype=np.float64), pid4, n_sites, seed, device).cpu().numpy()
    obs = (((st.astype(np.int32) % 2) - 1) * -1) + 1
    return atree.with_states(obs[0] if n_sites == 1 else obs.astype(np.uint8), segments=segments)
