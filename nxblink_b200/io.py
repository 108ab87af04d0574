"""File-format front end of the p53 example (replaces examples/p53/app_helper.py:44-184
and its dendropy dependency; the newick reader lives in nxblink_b200.p53.read_newick).

    read_phylip                      app_helper.py:83-107   (taxon name, codons) pairs
    read_genetic_code                app_helper.py:158-184  (state, residue, codon) triples
    read_interpreted_disease_data    app_helper.py:44-65    (position, residue, status) rows
    AlignmentData                    examples/p53/p53data.py:8-94  per-column data object
    pack_columns                     the same constraints for all columns at once (bit masks)
"""
from __future__ import division, print_function

import numpy as np

from .packing import pack_alignment

BENIGN, LETHAL, UNKNOWN = 'BENIGN', 'LETHAL', 'UNKNOWN'


def gen_paragraphs(lines):
    para = []
    for line in lines:
        line = line.strip()
        if line:
            para.append(line)
        elif para:
            yield para
            para = []
    if para:
        yield para


def read_phylip(fin, ntaxa=None, ncodons=None):
    """Yield (taxon name, list of codons); the first paragraph is the header."""
    paras = list(gen_paragraphs(fin))[1:]
    if ntaxa is not None and len(paras) != ntaxa:
        raise Exception('expected an alignment of %d taxa' % ntaxa)
    for para in paras:
        codons = ' '.join(para[1:]).split()
        if ncodons is not None and len(codons) != ncodons:
            raise Exception('expected %d codons' % ncodons)
        yield para[0], codons


def read_genetic_code(fin):
    code = []
    for line in fin:
        line = line.strip()
        if line:
            state, residue, codon = line.split()
            if residue.upper() != 'STOP':
                code.append((int(state), residue.upper(), codon.upper()))
    return code


def read_interpreted_disease_data(fin):
    rows = []
    for line in fin.readlines()[1:]:
        if line.strip():
            pos, residue, status = line.split()
            rows.append((int(pos), residue, status))
    return rows


class AlignmentData(object):
    """One codon column: the three getters of nxblink/raoteh.py:45-52 (p53data.py:8-94)."""

    def __init__(self, genetic_code, benign_residues, lethal_residues, all_nodes, names,
            name_to_leaf, reference_leaf, primary_to_tol, codon_to_state, codon_column):
        self._nodes = list(all_nodes)
        self._names, self._name_to_leaf = names, name_to_leaf
        self._ref, self._p2t, self._c2s = reference_leaf, primary_to_tol, codon_to_state
        self._column = codon_column
        self._benign, self._lethal = set(), set()
        for s, r, c in genetic_code:
            if r in benign_residues:
                self._benign.add(s)
            elif r in lethal_residues:
                self._lethal.add(s)
            else:
                raise Exception('residue %s is neither benign nor lethal' % r)

    def get_primary_data(self):
        allp = set(self._p2t)
        out = dict((v, set(allp)) for v in self._nodes)
        for name, codon in zip(self._names, self._column):
            out[self._name_to_leaf[name]] = {self._c2s[codon]}
        return out

    def get_tolerance_data(self):
        out = {}
        for part in set(self._p2t.values()):
            out[part] = dict((v, {False, True}) for v in self._nodes)
            for name, codon in zip(self._names, self._column):
                if self._p2t[self._c2s[codon]] == part:
                    out[part][self._name_to_leaf[name]] = {True}
        for s in self._benign:
            part = self._p2t[s]
            out[part][self._ref] = {True} & out[part][self._ref]
        for s in self._lethal:
            part = self._p2t[s]
            out[part][self._ref] = {False} & out[part][self._ref]
        return out

    def get_data(self):
        data = {'PRIMARY': self.get_primary_data()}
        data.update(self.get_tolerance_data())
        return data


def pack_columns(pm, model, names, codon_sequences, reference_name=None, disease_rows=None):
    """All codon columns of an alignment as one PackedData (run-stochastic.py:86-186 without
    the per-column Python objects).  ``disease_rows``: (position from 1, residue, status)."""
    S = len(codon_sequences[0])
    N = pm.n_nodes
    leaf_states = -np.ones((S, N), dtype=np.int64)
    for name, codons in zip(names, codon_sequences):
        v = pm.node_index[model.name_to_leaf[name]]
        leaf_states[:, v] = [pm.state_index[model.codon_to_state[c.upper()]] for c in codons]
    K = pm.n_classes
    allk = np.uint32((1 << K) - 1)
    tt = np.full((S, N), allk, dtype=np.uint32)
    tf = np.full((S, N), allk, dtype=np.uint32)
    if disease_rows:
        ref = pm.node_index[model.name_to_leaf[reference_name]]
        for pos, residue, status in disease_rows:
            k = pm.tol_index[model.residue_to_part[residue.upper()]]
            if status == BENIGN:
                tf[pos - 1, ref] &= ~np.uint32(1 << k)
            elif status == LETHAL:
                tt[pos - 1, ref] &= ~np.uint32(1 << k)
            else:
                raise NotImplementedError('unknown amino acid status')
    return pack_alignment(pm, leaf_states, tt, tf)
