"""Ingestion (sgvamp-py_b200/ingest.py) against what the UNMODIFIED reference driver hands to its solver
(tests/golden/ingest/reference.npz, made by tests/golden/make_ingest_golden.py from src/main.py run with
one thread per MPI rank): merged marker order, reordered XTy (.assoc.linear with NaN and sqrt(N) scaling),
PLINK .ld matrices after the exchange of missing SNPs, the regularised Rused, and the saved .bim."""
import os

import numpy as np
import pytest

import ingest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
D = os.path.join(GOLD, "ingest")
p = lambda n: os.path.join(D, n)


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(D, "reference.npz"), allow_pickle=False)


def test_k2_ld_bim_exchange_matches_reference(ref, tmp_path):
    M, Rs, rs, mg = ingest.load_all([p("c1.ld"), p("c2.ld")], [p("c1.assoc.linear"), p("c2.assoc.linear")],
                                    [p("c1.bim"), p("c2.bim")], [400, 900], [8, 9])
    assert M == int(ref["k2_M"]) == 10
    s = 0.2
    for k in range(2):
        Rused = (1 - s) * Rs[k].toarray() + s * np.eye(M)                  # src/main.py:265
        assert np.array_equal(Rused, ref["k2_R_%d" % k]), k
        assert np.array_equal(rs[k], ref["k2_r_%d" % k]), k
    out = tmp_path / "m.bim"
    ingest.write_ref_bim(mg["ref_df"], str(out))
    assert out.read_text() == str(ref["k2_bim"])
    # the reference picks the supplier of a missing SNP by argmax POSITION (src/main.py:162): with K = 2 that is
    # always 0, so cohort 0 never asks anybody (QUIRK) while cohort 1 asks cohort 0
    assert set(np.unique(mg["sources"][0])) == {0.0} and set(np.unique(mg["sources"][1])) == {0.0, 1.0}
    # the evident intention instead: both cohorts get their missing SNPs filled in
    M2, Rs2, rs2, mg2 = ingest.load_all([p("c1.ld"), p("c2.ld")], [p("c1.assoc.linear"), p("c2.assoc.linear")],
                                        [p("c1.bim"), p("c2.bim")], [400, 900], [8, 9], source_quirk=False)
    miss0 = [mg2["idx"][rs_] for rs_ in ("rs3", "rs7")]
    assert np.all(rs[0][miss0] == 0) and np.all(rs2[0][miss0] != 0)
    assert Rs2[0].nnz > Rs[0].nnz and np.array_equal(Rs2[1].toarray(), Rs[1].toarray())


def test_k3_merge_and_supplier_choice(tmp_path):
    """K = 3 is beyond what the reference driver can run with the installed pandas (its second `.bim` merge raises on
    duplicate `_y` columns), so there is no golden: the merge is checked on its definition (outer union, coordinate
    order, SNPs unknown to cohort 1 last with NaN coordinate exactly as for K = 2) and the supplier choice against a
    direct restatement of src/main.py:156-163."""
    rng = np.random.default_rng(7)
    snps = ["rs%d" % i for i in range(101, 113)]
    coord = {rs: 5000 + 11 * i for i, rs in enumerate(snps)}
    drop = [("rs103", "rs108"), ("rs105",), ("rs103", "rs110", "rs112")]
    N_list, lists = [400, 900, 600], []
    for k in range(3):
        ss = [rs for rs in snps if rs not in drop[k]]
        if k == 1:
            ss = ss[3:] + ss[:3]
        lists.append(ss)
        with open(tmp_path / ("d%d.bim" % k), "w") as f:
            for rs in ss:
                f.write("2\t%s\t0\t%d\tC\tT\n" % (rs, coord[rs]))
    mg = ingest.merge_bims([str(tmp_path / ("d%d.bim" % k)) for k in range(3)], N_list)
    assert mg["M"] == 12 and set(mg["ref"]) == set(snps)
    known = [rs for rs in mg["ref"] if rs in lists[0]]
    assert known == sorted(known, key=lambda r: coord[r])             # cohort 1's SNPs in coordinate order ...
    assert mg["ref"][len(known):] == [rs for rs in mg["ref"] if rs not in lists[0]]   # ... the others after them
    assert list(mg["ref_df"].columns) == ingest.BIM_COLUMNS
    for k in range(3):
        assert [mg["ref"][i] for i in mg["i_maps"][k]] == lists[k]
        src = np.ones(12) * k
        for rs in set(snps) - set(lists[k]):
            cand = [q for q in range(3) if q != k and rs in lists[q]]
            src[mg["idx"][rs]] = int(np.argmax(np.array(N_list)[cand]))                 # the reference's rule (position)
        assert np.array_equal(mg["sources"][k], src)
    fixed = ingest.merge_bims([str(tmp_path / ("d%d.bim" % k)) for k in range(3)], N_list, source_quirk=False)
    for k in range(3):
        for rs in set(snps) - set(lists[k]):
            q = int(fixed["sources"][k][fixed["idx"][rs]])
            assert q != k and rs in lists[q]                                             # a cohort that really has the SNP
            assert N_list[q] == max(N_list[c] for c in range(3) if c != k and rs in lists[c])


def test_k1_ld_matches_reference(ref, tmp_path):
    M, Rs, rs, mg = ingest.load_all([p("c2.ld")], [p("c2.assoc.linear")], [p("c2.bim")], [900], [9])
    assert M == int(ref["k1_M"]) == 9
    assert np.array_equal(Rs[0].toarray(), ref["k1_R_0"])
    assert np.array_equal(rs[0], ref["k1_r_0"])
    out = tmp_path / "m.bim"
    ingest.write_ref_bim(mg["ref_df"], str(out))
    assert out.read_text() == str(ref["k1_bim"])
    R = Rs[0]
    assert (R != R.T).nnz == 0 and np.all(R.diagonal() == 1.0)


def test_readers_and_errors(tmp_path):
    np.save(tmp_path / "r.npy", np.arange(5.0))
    np.savetxt(tmp_path / "r.txt", np.arange(5.0))
    assert np.array_equal(ingest.load_r(str(tmp_path / "r.npy"), 5, 10), np.arange(5.0))
    assert np.array_equal(ingest.load_r(str(tmp_path / "r.txt"), 5, 10), np.arange(5.0))
    with pytest.raises(Exception, match="Unsupported r vector format"):
        ingest.load_r("r.csv", 5, 10)
    with pytest.raises(Exception, match="Unsupported R matrix format"):
        ingest.load_R_file("R.mtx")
    with pytest.raises(Exception, match="needs --bim-files"):
        ingest.load_all(["a.ld"], ["r.npy"], None, [10], [5])
    with pytest.raises(Exception, match="different marker sets"):
        ingest.load_all(["a.npz", "b.npz"], ["r.npy", "q.npy"], None, [10, 10], [5, 6])
