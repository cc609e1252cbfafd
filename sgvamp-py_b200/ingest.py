"""Input ingestion of the command-line driver: the `.bim` reference-order merge, the XTy readers, the
PLINK `.ld` reader and the exchange of missing SNPs between cohorts (reference src/main.py:126-260).

The reference runs one MPI rank per cohort and exchanges the rows of missing SNPs with
`comm.send/recv` (src/main.py:211-249).  Here one process holds all K cohorts, so the exchange is a
lookup in memory; everything else follows the reference line by line, including its observable
quirks (kept so that outputs match; each is marked QUIRK):

  * the reference marker order is the outer merge of the K `.bim` tables on `Variant`, sorted by
    `Coordinate` (src/main.py:130-143);
  * QUIRK `source` (src/main.py:156-163): for a SNP missing in cohort k the supplying cohort is chosen
    as `np.argmax(N_list[candidates])`, which is the POSITION inside the candidate list, not the cohort
    number.  `source_quirk=False` selects the evident intention (the candidate with the largest N);
  * QUIRK duplicates (src/main.py:228-233, 257-258): a supplier sends every `.ld` row that touches a
    requested SNP once per requested SNP it touches, and `csr_matrix((v, (i, j)))` adds duplicates.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse

BIM_COLUMNS = ["Chromosome", "Variant", "Position", "Coordinate", "Allele1", "Allele2"]


def read_bim(path):
    import pandas as pd
    return pd.read_table(path, sep=r"\s+", header=None, names=BIM_COLUMNS)           # src/main.py:133


def merge_bims(bim_paths, N_list, source_quirk=True):
    """Reference marker order and the per-cohort index maps (src/main.py:130-164)."""
    import pandas as pd
    K = len(bim_paths)
    lists, ref_df = [], None
    for k in range(K):
        df = read_bim(bim_paths[k])
        lists.append(list(df["Variant"]))
        if k == 0:
            ref_df = df
        else:
            ref_df = pd.merge(ref_df, df, on=["Variant"], how="outer", suffixes=("", "_y"))               # :138
            # Only the first six columns are ever used (:140,:150).  With K >= 3 the reference's second merge raises
            # (pandas refuses the duplicate `_y` columns), i.e. the reference itself stops at K = 2 here; dropping the
            # unused `_y` columns after every merge gives the same result for K = 2 and extends it to any K.
            ref_df = ref_df[[c for c in ref_df.columns if not c.endswith("_y")]]
    ref_df = ref_df.sort_values(by=["Coordinate"])                                                       # :140
    ref = list(ref_df["Variant"])
    M = len(ref)
    idx = {rs: i for i, rs in enumerate(ref)}
    sets = [set(x) for x in lists]
    i_maps, sources = [], []
    for k in range(K):
        i_maps.append(np.array([idx[rs] for rs in lists[k]], dtype=np.int64))                           # :153
        source = np.ones(M) * k                                                                           # :155
        for rs in set(ref) - sets[k]:
            cand = [q for q in range(K) if q != k and rs in sets[q]]
            kx = int(np.argmax(np.array(N_list)[cand]))                                                   # :162
            source[idx[rs]] = kx if source_quirk else cand[kx]                                            # QUIRK: position, not cohort
        sources.append(source)
    return dict(ref_df=ref_df, ref=ref, M=M, idx=idx, lists=lists, i_maps=i_maps, sources=sources)


def write_ref_bim(ref_df, path):
    ref_df.iloc[:, :6].to_csv(path, header=None, sep="\t", index=False)                                  # src/main.py:150


def load_r(path, M_k, N_k):
    """XTy of one cohort in its own marker order (src/main.py:176-187)."""
    if path.endswith(".txt"):
        return np.loadtxt(path).reshape(M_k).astype(np.float64)
    if path.endswith(".npy"):
        return np.load(path).reshape(M_k).astype(np.float64)
    if path.endswith(".linear"):
        import pandas as pd
        df = pd.read_table(path, sep=r"\s+")
        r = np.array(df["BETA"], dtype=np.float64).reshape(M_k)
        r[np.isnan(r)] = 0
        return r * np.sqrt(N_k)                                                                           # :185
    raise Exception("Unsupported r vector format!")


def reorder_r(r_k, i_map, M):
    r = np.zeros(M)
    r[i_map] = r_k                                                                                        # src/main.py:190-191
    return r


def load_ld_triples(path, idx):
    """PLINK `.ld` table -> (indA, indB, R) in reference indices (src/main.py:205-208)."""
    import pandas as pd
    df = pd.read_table(path, sep=r"\s+")
    a = np.array([idx[rs] for rs in df["SNP_A"]], dtype=np.int64)
    b = np.array([idx[rs] for rs in df["SNP_B"]], dtype=np.int64)
    return a, b, np.array(df["R"], dtype=np.float64)


def exchange_missing(sources, triples, rs):
    """What the send/recv rounds of src/main.py:211-249 leave on every cohort: its own `.ld` rows plus the
    rows received for the SNPs it asked other cohorts for, and r filled in at those SNPs.
    triples[k] = (indA, indB, R) or None for a cohort whose LD did not come from a `.ld` file."""
    K = len(sources)
    out_t, out_r = [], []
    for k in range(K):
        if triples[k] is None:
            out_t.append(None)
            out_r.append(rs[k])
            continue
        A, B, V = [triples[k][0]], [triples[k][1]], [triples[k][2]]
        r = rs[k].copy()
        for q in range(K):                                                 # receive loop :236-249
            if q == k or not (sources[k] == q).any():
                continue
            req = np.flatnonzero(sources[k] == q)                          # request list :213
            if triples[q] is None:
                raise Exception("cohort %d asks cohort %d for missing SNPs, but its LD is not a .ld file" % (k, q))
            qa, qb, qv = triples[q]
            for ind in req:                                                # supplier side :226-231
                hit = (qa == ind) | (qb == ind)
                A.append(qa[hit]); B.append(qb[hit]); V.append(qv[hit])   # QUIRK: once per requested SNP touched
            r[sources[k] == q] = rs[q][req]                                # :248 (the supplier's original r, :232)
        out_t.append((np.concatenate(A), np.concatenate(B), np.concatenate(V)))
        out_r.append(r)
    return out_t, out_r


def build_R(M, indA, indB, vals):
    """Symmetric CSR with unit diagonal; duplicate entries add up (src/main.py:251-258)."""
    ind_r = np.concatenate([np.arange(M), indA, indB])
    ind_c = np.concatenate([np.arange(M), indB, indA])
    v = np.concatenate([np.ones(M), vals, vals])
    return scipy.sparse.csr_matrix((v, (ind_r, ind_c)), shape=(M, M))


def load_R_file(path):
    if path.endswith(".npz"):
        return scipy.sparse.load_npz(path)
    if path.endswith(".npy"):
        return np.load(path, mmap_mode="r")
    raise Exception("Unsupported R matrix format!")


def load_all(ld_paths, r_paths, bim_paths, N_list, M_list, source_quirk=True):
    """Everything src/main.py does between argument parsing and `Rused`: returns (M, R_list, r_list, merge)
    with every cohort in the reference marker order.  Without `.bim` files the cohorts must share one
    marker order (the reference cannot run without them, src/main.py:81)."""
    K = len(ld_paths)
    if bim_paths is None:
        if len(set(M_list)) != 1:
            raise Exception("cohorts with different marker sets need --bim-files")
        if any(p.endswith(".ld") for p in ld_paths):
            raise Exception("PLINK .ld input needs --bim-files (SNP names are resolved through the .bim tables)")
        M = M_list[0]
        return M, [load_R_file(p) for p in ld_paths], [load_r(r_paths[k], M, N_list[k]) for k in range(K)], None
    mg = merge_bims(bim_paths, N_list, source_quirk)
    M = mg["M"]
    rs = [reorder_r(load_r(r_paths[k], M_list[k], N_list[k]), mg["i_maps"][k], M) for k in range(K)]
    triples = [load_ld_triples(p, mg["idx"]) if p.endswith(".ld") else None for p in ld_paths]
    triples, rs = exchange_missing(mg["sources"], triples, rs)
    Rs = []
    for k in range(K):
        if triples[k] is not None:
            Rs.append(build_R(M, *triples[k]))
        else:
            R = load_R_file(ld_paths[k])
            if R.shape != (M, M):
                raise Exception("LD matrix %s has shape %s but the merged marker set has M=%d; .npz/.npy LD must be "
                                "in the reference order of all markers" % (ld_paths[k], R.shape, M))
            Rs.append(R)
    return M, Rs, rs, mg
