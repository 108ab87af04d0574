"""File-format front end (app_helper.py / p53data.py mirror): the vectorised column packing
equals the per-column data objects; the readers parse the reference's formats."""
import io

import numpy as np

from nxblink_b200 import io as nio, p53, packing


def _toy_alignment(model, pm, ncols=5, seed=0):
    rng = np.random.default_rng(seed)
    names = sorted(model.name_to_leaf)
    codons = [c for s, r, c in model.genetic_code]
    seqs = [[codons[i] for i in rng.integers(0, 61, size=ncols)] for _ in names]
    return names, seqs


def test_readers_parse_reference_formats():
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    names, seqs = _toy_alignment(m, pm)
    text = '%d %d\n\n' % (len(names), len(seqs[0]))
    for n, sq in zip(names, seqs):
        text += n + '\n' + ' '.join(sq[:3]) + '\n' + ' '.join(sq[3:]) + '\n\n'
    got = list(nio.read_phylip(io.StringIO(text), ntaxa=25, ncodons=5))
    assert [g[0] for g in got] == names and [g[1] for g in got] == seqs
    code_text = ''.join('%d\t%s\t%s\n' % (s, r.lower(), c.lower()) for s, r, c in m.genetic_code)
    code_text += '61\tstop\ttaa\n'
    assert nio.read_genetic_code(io.StringIO(code_text)) == m.genetic_code
    dis = 'position\tresidue\tstatus\n1\tALA\tBENIGN\n1\tARG\tLETHAL\n\n'
    assert nio.read_interpreted_disease_data(io.StringIO(dis)) == [(1, 'ALA', 'BENIGN'), (1, 'ARG', 'LETHAL')]


def test_pack_columns_equals_per_column_objects():
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    names, seqs = _toy_alignment(m, pm, ncols=4, seed=3)
    residues = sorted(set(r for s, r, c in m.genetic_code))
    rng = np.random.default_rng(1)
    rows, per_col = [], []
    ref_col = seqs[names.index('Has')]
    for pos in range(1, 5):
        own = [r for s, r, c in m.genetic_code if c == ref_col[pos - 1]][0]
        benign, lethal = set(), set()
        for r in residues:
            status = nio.BENIGN if (r == own or rng.random() > 0.3) else nio.LETHAL
            rows.append((pos, r, status))
            (benign if status == nio.BENIGN else lethal).add(r)
        per_col.append((benign, lethal))
    fast = nio.pack_columns(pm, m, names, seqs, reference_name='Has', disease_rows=rows)
    objs = []
    T, root = m.get_T_and_root()
    for i, (benign, lethal) in enumerate(per_col):
        objs.append(nio.AlignmentData(m.genetic_code, benign, lethal, list(T), names, m.name_to_leaf,
            m.name_to_leaf['Has'], m.get_primary_to_tol(), m.codon_to_state,
            [sq[i] for sq in seqs]))
    slow = packing.pack_data(pm, objs)
    assert (fast.pri_ok == slow.pri_ok).all()
    assert (fast.tol_true_ok == slow.tol_true_ok).all()
    assert (fast.tol_false_ok == slow.tol_false_ok).all()
