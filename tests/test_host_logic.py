"""CPU tests: host-side label logic, the C-ABI library surface, API argument checks."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from oracle import mn_oracle as O
from pymn_b200 import _lib
from pymn_b200.utils import join_labels, joint_layout, label_layout
from tests.helpers import ROOT, make_adata
from oracle import cases as C


def test_label_layout_matches_reference_order():
    # "s1" vs "s10": '|' sorts after digits, so the joined-string order differs from the
    # per-study order
    study = np.array(["s10", "s1", "s2", "s1", "s10", "s2", "s1"], dtype=object)
    ctype = np.array(["b", "a", "a", "b", "a", "c", "a"], dtype=object)
    lay = label_layout(study, ctype)
    assert list(lay.studies) == ["s1", "s10", "s2"]
    assert list(lay.row_labels) == ["s1|a", "s1|b", "s10|a", "s10|b", "s2|a", "s2|c"]
    # global order = sorted joined strings
    glob = [lay.row_labels[i] for i in np.argsort(lay.sorted_pos)]
    assert glob == sorted(lay.row_labels.tolist())
    assert glob[0] == "s10|a"
    # device rows are grouped by label, labels by study
    assert np.all(np.diff(lay.cell_label) >= 0)
    assert list(lay.seg_off) == [0, 3, 5, 7]
    assert list(lay.seg_lab_off) == [0, 2, 4, 6]
    joined = join_labels(study, ctype)
    assert [lay.row_labels[c] for c in lay.cell_label] == list(joined[lay.perm])
    # agrees with the oracle's encoding
    uniq, _ = O.encode(O.join_labels(study, ctype))
    assert list(uniq) == glob


def test_joint_layout_has_all_combinations():
    study = np.array(["a", "a", "b", "b", "b"], dtype=object)
    ctype = np.array(["x", "y", "y", "y", "z"], dtype=object)
    lay, names = joint_layout(study, ctype)
    assert list(names) == ["x", "y", "z"]
    assert len(lay.row_labels) == 6
    assert list(lay.lab_row_off) == [0, 1, 2, 2, 2, 4, 5]
    assert list(lay.seg_off) == [0, 2, 5]


def test_join_labels_replace_bar_warns():
    with pytest.warns(UserWarning):
        out = join_labels(np.array(["a|b"], dtype=object), np.array(["t"], dtype=object), replace_bar=True)
    assert out[0] == "a.b|t"


def test_shared_library_exports_every_declared_symbol():
    """include/pymn_b200.h, the ctypes table and the built library agree (no compute calls)."""
    header = open(os.path.join(ROOT, "include", "pymn_b200.h")).read()
    declared = set(re.findall(r"\b(mn_[a-z0-9_]+)\s*\(", header))
    declared -= {"mn_stream_t", "mn_half_t"}
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert os.path.isfile(_lib.LIB_PATH), "build the library first: pymn_b200/csrc/build.sh"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert _lib.load().mn_abi_version() == 3


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pymn_b200
    spec, inp, _ = C.load_case("us_fast_ndn")
    ad = make_adata(inp)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", fast_version=True, save_uns=False)


def test_api_assertions_match_reference_messages():
    import pymn_b200
    spec, inp, _ = C.load_case("us_fast_ndn")
    ad = make_adata(inp)
    with pytest.raises(AssertionError, match="Study Col not in adata"):
        pymn_b200.MetaNeighborUS(ad, "nope", "cell_type")
    with pytest.raises(AssertionError, match="Cluster Col not in adata"):
        pymn_b200.MetaNeighborUS(ad, "study_id", "nope")
    with pytest.raises(AssertionError, match="must be in adata.var_keys"):
        pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", var_genes="missing")
    with pytest.raises(AssertionError, match="one_vs_best"):
        pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", one_vs_best=True)
    ad.obs["study_id"] = ad.obs["study_id"].astype("category")
    with pytest.raises(AssertionError, match="category type"):
        pymn_b200.MetaNeighborUS(ad, "study_id", "cell_type", fast_version=True)
    one = make_adata(inp)
    one.obs["study_id"] = "only"
    with pytest.raises(AssertionError, match="Need more than 1 study"):
        pymn_b200.MetaNeighborUS(one, "study_id", "cell_type", fast_version=True)
    with pytest.raises(AssertionError, match="Study Col not in adata"):
        pymn_b200.trainModel(ad, "nope", "cell_type")
    gs = pd.DataFrame(np.ones((3, 1), int), index=["zz1", "zz2", "zz3"], columns=["s"])
    with pytest.raises(AssertionError, match="No matching genes"):
        pymn_b200.MetaNeighbor(make_adata(inp), "study_id", "cell_type", gs)


def test_shard_and_gather_single_rank():
    from pymn_b200.MetaNeighbor import gather_columns, shard_columns
    assert list(shard_columns(7, 1, 3)) == [1, 4]
    a = np.arange(6.0).reshape(2, 3)
    assert gather_columns(a, 3, 0, 1) is a


def test_model_csv_round_trip(tmp_path):
    """the (1 + G) x L model table survives to_csv / read_csv(index_col=0) bit for bit,
    a file written R-style (quoted header, empty corner cell) loads too, and bad files fail early."""
    import pandas as pd
    import pymn_b200
    rng = np.random.default_rng(0)
    genes = [f"g{i}" for i in range(25)]
    cols = ["s1|a", "s1|b c", "s10|a"]
    vals = np.vstack([[12, 7, 30], rng.normal(size=(25, 3))])
    model = pd.DataFrame(vals, index=["n_cells"] + genes, columns=cols)
    path = tmp_path / "model.csv"
    pymn_b200.save_model(model, path)
    back = pymn_b200.load_model(path)
    assert list(back.index) == list(model.index) and list(back.columns) == cols
    assert np.array_equal(back.values, model.values)
    r_path = tmp_path / "model_r.csv"
    with open(r_path, "w") as f:
        f.write('"",' + ",".join(f'"{c}"' for c in cols) + "\n")
        for name, row in zip(model.index, model.values):
            f.write(f'"{name}",' + ",".join(repr(float(v)) for v in row) + "\n")
    assert np.array_equal(pymn_b200.load_model(r_path).values, model.values)
    bad = model.copy()
    bad.index = ["cells"] + genes
    with pytest.raises(AssertionError):
        pymn_b200.save_model(bad, tmp_path / "bad.csv")


def test_select_genes_identity_is_free():
    """Selecting every gene in order must not copy the matrix; any other selection slices as the reference does."""
    from pymn_b200.utils import select_genes
    rng = np.random.default_rng(1)
    x = rng.poisson(1.0, size=(20, 6)).astype(np.float32)
    inp = dict(X=x, study=np.array(["a"] * 20, dtype=object), ctype=np.array(["t"] * 20, dtype=object),
               genes=np.array([f"g{i}" for i in range(6)], dtype=object), hvg=np.ones(6, bool))
    ad = make_adata(inp)
    assert select_genes(ad, ad.var_names) is ad
    assert select_genes(ad, ad.var_names[np.ones(6, bool)]) is ad
    sub = select_genes(ad, ad.var_names[[0, 2, 5]])
    assert sub is not ad and np.array_equal(sub.X, x[:, [0, 2, 5]])
    rev = select_genes(ad, ad.var_names[::-1])
    assert rev is not ad and np.array_equal(rev.X, x[:, ::-1])


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("kind", ["object", "arrow", "int", "bar"])
def test_label_layout_randomised_against_plain_python(seed, kind):
    """label_layout (Arrow / hash factorisation, O(N) pair coding, radix argsort) against a direct restatement
    of what the reference builds: joined 'study|type' strings, np.unique order, per-study sorted rows."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    st_pool = ["s1", "s10", "s2", "S", "a b", "z|z" if kind == "bar" else "zz"][: int(rng.integers(1, 6))]
    ct_pool = ["t1", "t10", "T", "x y", "é", "t2", "a"][: int(rng.integers(1, 7))]
    if kind == "int":
        study = rng.integers(0, 12, size=n)
        ctype = rng.integers(-3, 20, size=n)
    else:
        study = np.array([st_pool[i] for i in rng.integers(0, len(st_pool), size=n)], dtype=object)
        ctype = np.array([ct_pool[i] for i in rng.integers(0, len(ct_pool), size=n)], dtype=object)
    s_in, c_in = study, ctype
    if kind == "arrow":
        df = pd.DataFrame({"s": study, "c": ctype})
        s_in, c_in = df["s"].values, df["c"].values            # pandas string arrays (Arrow-backed on pandas 3)
    replace_bar = kind == "bar"
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lay = label_layout(s_in, c_in, replace_bar=replace_bar)
    s_str = [str(v).replace("|", ".") if replace_bar else str(v) for v in study.tolist()]
    c_str = [str(v) for v in ctype.tolist()]
    joined = [a + "|" + b for a, b in zip(s_str, c_str)]
    if kind == "int":
        st_sorted = [str(v) for v in np.unique(study).tolist()]           # numeric order, as np.unique on ints
    else:
        st_sorted = sorted(set(str(v) for v in study.tolist()))           # studies sorted BEFORE the bar replacement
        st_sorted = [v.replace("|", ".") if replace_bar else v for v in st_sorted]
    rows = []
    for st in st_sorted:
        rows += sorted({j for j, s in zip(joined, s_str) if s == st})
    assert list(lay.row_labels) == rows
    assert [lay.row_labels[c] for c in lay.cell_label] == [joined[i] for i in lay.perm]
    assert sorted(lay.perm.tolist()) == list(range(n))
    # stable within a label: input order preserved
    for lab in range(len(rows)):
        idx = lay.perm[lay.lab_row_off[lab]:lay.lab_row_off[lab + 1]]
        assert np.all(np.diff(idx) > 0)
    assert [lay.row_labels[i] for i in np.argsort(lay.sorted_pos)] == sorted(rows)
    assert lay.seg_off[-1] == n and lay.lab_row_off[-1] == n
    for t in range(len(st_sorted)):
        assert set(lay.cell_study[lay.seg_off[t]:lay.seg_off[t + 1]].tolist()) <= {t}


def test_gene_count_limit_is_reported_early():
    """More than 16,384 selected genes: a clear ValueError at the API instead of a failure deep inside the library."""
    import pymn_b200
    from pymn_b200.anndata_lite import AnnDataLite
    G = 16385
    ad = AnnDataLite(np.zeros((4, G), np.float32), pd.DataFrame({"s": ["a", "a", "b", "b"], "c": ["x", "y", "x", "y"]}),
                     pd.DataFrame({"highly_variable": np.ones(G, bool)}, index=[f"g{i}" for i in range(G)]))
    with pytest.raises(ValueError, match="16384"):
        pymn_b200.MetaNeighborUS(ad, "s", "c", fast_version=True)
    with pytest.raises(ValueError, match="16384"):
        pymn_b200.trainModel(ad, "s", "c")


def test_gc_paused_restores_collector_state():
    """The public entry points pause the cyclic GC for the duration of a call and restore whatever state they found."""
    import gc
    from pymn_b200.utils import gc_paused, with_gc_paused
    was = gc.isenabled()
    try:
        gc.enable()
        with gc_paused():
            assert not gc.isenabled()
        assert gc.isenabled()
        gc.disable()
        with gc_paused():
            assert not gc.isenabled()
        assert not gc.isenabled()                     # it was off before: stays off

        @with_gc_paused
        def f(a, b=2):
            """doc"""
            assert not gc.isenabled()
            raise ValueError("boom")
        gc.enable()
        with pytest.raises(ValueError):
            f(1)
        assert gc.isenabled() and f.__doc__ == "doc"
    finally:
        (gc.enable if was else gc.disable)()


def test_label_layout_helper_thread_and_sequential_agree(monkeypatch):
    """MN_LABEL_HELPER: encoding the second obs column on the helper thread gives the layout of the sequential default."""
    import pymn_b200.utils as U
    rng = np.random.default_rng(5)
    n = 60_000
    df = pd.DataFrame({"s": np.array([f"s{v}" for v in rng.integers(0, 4, size=n)], dtype=object),
                       "c": np.array([f"c{v}" for v in rng.integers(0, 30, size=n)], dtype=object)})
    outs = []
    for helper in (False, True):
        monkeypatch.setattr(U, "LABEL_HELPER", helper)
        outs.append(label_layout(df["s"].values, df["c"].values))
    for f in ("perm", "cell_label", "cell_study", "seg_off", "lab_row_off", "seg_lab_off", "label_study", "sorted_pos"):
        assert np.array_equal(getattr(outs[0], f), getattr(outs[1], f)), f
    assert list(outs[0].row_labels) == list(outs[1].row_labels)
