"""The p53-shaped codon blink model: MG94 primary process, 20 amino-acid
tolerance classes, the 25-taxon p53 tree, and synthetic alignments of that shape.

Mirrors, for model construction only:
    examples/p53/create_mg94.py:60-178   MG94 rate matrix + codon distribution
    examples/p53/p53model.py:55-66       parameter values (kappa, nucleotide freqs)
    examples/p53/p53model.py:143-232     Model (the nine getters)
    examples/p53/run-stochastic.py:50-66,131-141   blen := 1, edge rate := branch length,
                                                   rate_on = 1.5, rate_off = 0.5
    examples/p53/p53data.py:52-94        alignment + disease constraints
The newick reader replaces the reference's dendropy dependency
(examples/p53/app_helper.py:100-137).
"""
from __future__ import division, print_function

import os

import numpy as np

from .summary import WeightGraph
from .toymodel import Tree

_HERE = os.path.dirname(os.path.abspath(__file__))
P53_TREE_PATH = os.path.join(_HERE, 'data', 'p53_tree.newick')

_NT = 'TCAG'
_AA = 'FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG'
_AA3 = dict(A='ALA', R='ARG', N='ASN', D='ASP', C='CYS', Q='GLN', E='GLU', G='GLY',
        H='HIS', I='ILE', L='LEU', K='LYS', M='MET', F='PHE', P='PRO', S='SER',
        T='THR', W='TRP', Y='TYR', V='VAL')


def universal_genetic_code():
    """[(state, residue, codon)] for the 61 sense codons, ordered like
    examples/p53/universal.code.txt (residues alphabetical, codons in TCAG order)."""
    rows = []
    i = 0
    for a in _NT:
        for b in _NT:
            for c in _NT:
                aa = _AA[i]
                i += 1
                if aa != '*':
                    rows.append((_AA3[aa], i, a + b + c))
    rows.sort(key=lambda r: (r[0], r[1]))
    return [(s, res, codon) for s, (res, _, codon) in enumerate(rows)]


def jeff_params():
    """examples/p53/p53model.py:55-66 (omega fixed to 1: blinking replaces it)."""
    AG = 0.50862
    CT = 1 - AG
    A = AG * 0.49373
    G = AG - A
    T = CT * 0.38884
    C = CT - T
    return dict(kappa=3.38714, omega=1.0, A=A, C=C, G=G, T=T)


def create_mg94(A, C, G, T, kappa, omega, genetic_code, target_expected_rate=1.0):
    """Muse-Gaut 1994 codon model (create_mg94.py:60-178).

    Returns (Q, distn, state_to_residue, residue_to_part) like the reference.
    """
    nt = dict(A=A, C=C, G=G, T=T)
    transitions = ('AG', 'GA', 'CT', 'TC')
    residues = sorted(set(r for s, r, c in genetic_code))
    residue_to_part = dict((r, i) for i, r in enumerate(residues))
    state_to_residue = dict((s, r) for s, r, c in genetic_code)
    rates = {}
    for a, (sa, ra, ca) in enumerate(genetic_code):
        for b, (sb, rb, cb) in enumerate(genetic_code):
            diff = [(x, y) for x, y in zip(ca, cb) if x != y]
            if len(diff) != 1:
                continue
            x, y = diff[0]
            rate = nt[y]
            if x + y in transitions:
                rate *= kappa
            if ra != rb:
                rate *= omega
            rates[a, b] = rate
    w = np.array([np.prod([nt[n] for n in c]) for s, r, c in genetic_code])
    w = w / w.sum()
    distn = dict((i, float(w[i])) for i in range(len(genetic_code)))
    expected = sum(distn[a] * r for (a, b), r in rates.items())
    scale = target_expected_rate / expected
    Q = WeightGraph()
    for (a, b), r in sorted(rates.items()):
        Q.add_edge(a, b, r * scale)
    return Q, distn, state_to_residue, residue_to_part


def read_newick(text):
    """Minimal newick reader -> (Tree, root, edge_to_blen, name_to_leaf).

    Leaves are numbered first (in order of appearance), internal nodes after,
    the root last -- the numbering convention of app_helper.py:116-121.
    """
    text = text.strip().rstrip(';')
    pos = [0]
    leaves, internal = [], []

    def parse():
        node = {'children': [], 'name': None, 'length': None}
        if text[pos[0]] == '(':
            pos[0] += 1
            while True:
                node['children'].append(parse())
                if text[pos[0]] == ',':
                    pos[0] += 1
                    continue
                if text[pos[0]] == ')':
                    pos[0] += 1
                    break
                raise ValueError('bad newick at %d' % pos[0])
        start = pos[0]
        while pos[0] < len(text) and text[pos[0]] not in ':,()':
            pos[0] += 1
        name = text[start:pos[0]].strip()
        if name:
            node['name'] = name
        if pos[0] < len(text) and text[pos[0]] == ':':
            pos[0] += 1
            start = pos[0]
            while pos[0] < len(text) and text[pos[0]] not in ',()':
                pos[0] += 1
            node['length'] = float(text[start:pos[0]])
        (internal if node['children'] else leaves).append(node)
        return node

    root = parse()
    for i, n in enumerate(leaves + internal):
        n['id'] = i
    T = Tree()
    edge_to_blen = {}
    name_to_leaf = {}
    stack = [root]
    while stack:
        n = stack.pop()
        T.setdefault(n['id'], [])
        if n['name'] and not n['children']:
            name_to_leaf[n['name']] = n['id']
        for c in n['children']:
            T[n['id']].append(c['id'])
            edge_to_blen[n['id'], c['id']] = c['length'] if c['length'] is not None else 0.0
            stack.append(c)
    return T, root['id'], edge_to_blen, name_to_leaf


class Model(object):
    """The nine getters of nxblink/raoteh.py:33-44 (p53model.py:143-232)."""

    def __init__(self, kappa, omega, A, C, G, T, rate_on, rate_off,
            tree, root, edge_to_blen, edge_to_rate, genetic_code):
        self._rate_on, self._rate_off = rate_on, rate_off
        self._tree, self._root = tree, root
        self._edge_to_blen, self._edge_to_rate = edge_to_blen, edge_to_rate
        Q, distn, state_to_residue, residue_to_part = create_mg94(
                A, C, G, T, kappa, omega, genetic_code, target_expected_rate=1.0)
        self._Q, self._primary_distn = Q, distn
        self._primary_to_part = dict(
                (i, residue_to_part[r]) for i, r in state_to_residue.items())
        self.genetic_code = genetic_code
        self.residue_to_part = residue_to_part
        self.codon_to_state = dict((c, s) for s, r, c in genetic_code)

    def get_rate_on(self):
        return self._rate_on

    def get_rate_off(self):
        return self._rate_off

    def get_blink_distn(self):
        total = self._rate_on + self._rate_off
        return {0: self._rate_off / total, 1: self._rate_on / total}

    def get_primary_to_tol(self):
        return self._primary_to_part

    def get_T_and_root(self):
        return self._tree, self._root

    def get_edge_to_blen(self):
        return self._edge_to_blen

    def get_edge_to_rate(self):
        return self._edge_to_rate

    def get_primary_distn(self):
        return self._primary_distn

    def get_Q_primary(self):
        return self._Q


def p53_model(rate_on=1.5, rate_off=0.5, newick_path=None):
    """Config 3 of SURVEY.md section 8(d): blen := 1, edge rate := branch length."""
    with open(newick_path or P53_TREE_PATH) as fin:
        tree, root, branch_length, name_to_leaf = read_newick(fin.read())
    edge_to_rate = dict(branch_length)
    edge_to_blen = dict((e, 1.0) for e in edge_to_rate)
    prm = jeff_params()
    model = Model(prm['kappa'], prm['omega'], prm['A'], prm['C'], prm['G'], prm['T'],
            rate_on, rate_off, tree, root, edge_to_blen, edge_to_rate,
            universal_genetic_code())
    model.name_to_leaf = name_to_leaf
    return model


def synthetic_alignment(pm, n_sites, seed=0, reference_leaf=None, lethal_fraction=0.19):
    """Synthetic p53-shaped data for ``pack_alignment``.

    Leaf states are forward-simulated from the primary process along the tree
    (a plain CTMC with the model's normalised Q and per-edge rate * length; the
    tolerance process is ignored when generating data).  At ``reference_leaf``
    (internal node index) every class other than the observed one is declared
    LETHAL (tolerance forced OFF) with probability ``lethal_fraction`` and BENIGN
    (forced ON) otherwise, as in examples/p53/p53data.py:83-92; 1486 of the
    7860 rows of examples/p53/int1.out are LETHAL (0.19).

    Returns (leaf_states [S, N] with -1 at internal nodes, tol_true_ok, tol_false_ok).
    """
    import scipy.linalg
    rng = np.random.default_rng(seed)
    P, K, N = pm.n_states, pm.n_classes, pm.n_nodes
    Q = np.zeros((P, P))
    Q[pm.q_row, pm.q_col] = pm.q_val
    Q -= np.diag(Q.sum(axis=1))
    states = np.zeros((n_sites, N), dtype=np.int64)
    states[:, 0] = rng.choice(P, size=n_sites, p=pm.prior / pm.prior.sum())
    for v in range(1, N):
        Pt = scipy.linalg.expm(Q * (pm.blen[v] * pm.rate[v]))
        Pt = np.clip(Pt, 0, None)
        cdf = np.cumsum(Pt / Pt.sum(axis=1, keepdims=True), axis=1)
        u = rng.random(n_sites)
        states[:, v] = np.minimum(
                (u[:, None] > cdf[states[:, pm.parent[v]]]).sum(axis=1), P - 1)
    is_leaf = np.ones(N, dtype=bool)
    is_leaf[pm.parent[1:]] = False
    leaf_states = np.where(is_leaf[None, :], states, -1)
    allk = np.uint32((1 << K) - 1)
    tt = np.full((n_sites, N), allk, dtype=np.uint32)
    tf = np.full((n_sites, N), allk, dtype=np.uint32)
    if reference_leaf is not None:
        lethal = rng.random((n_sites, K)) < lethal_fraction
        own = pm.state_class[states[:, reference_leaf]]
        lethal[np.arange(n_sites), own] = False
        lethal_mask = (lethal * (1 << np.arange(K))).sum(axis=1).astype(np.uint32)
        tt[:, reference_leaf] = allk & ~lethal_mask      # lethal: ON infeasible
        tf[:, reference_leaf] = lethal_mask              # benign: OFF infeasible
    return leaf_states, tt, tf
