"""Shared driver for the golden fixtures (tests/golden/golden.npz, made by make_golden.py from the
real reference).  ``run_case(case, impl)`` replays one recorded reference call through ``impl`` —
the numpy oracle, the ATen port or the CUDA product — and yields (key, got, want) triples."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

_npz = None


def arrays():
    global _npz
    if _npz is None:
        _npz = np.load(os.path.join(GOLDEN, "golden.npz"))
    return _npz


def cases():
    return json.load(open(os.path.join(GOLDEN, "cases.json")))["cases"]


def ellipses():
    return json.load(open(os.path.join(GOLDEN, "ellipses.json")))


def case_inputs(case):
    z = arrays()
    pre = case["name"] + "/in/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def _flatten(out, to_np):
    flat = {}
    if isinstance(out, dict):
        for k, v in out.items():
            if v is None or k in ("xs", "ys", "covs", "sizes", "features"):
                continue
            if isinstance(v, dict):
                for kk, vv in v.items():
                    flat[f"{k}/{kk}"] = to_np(vv)
            else:
                flat[k] = to_np(v)
    else:
        flat["ret"] = to_np(out)
    return flat


def run_case(case, impl, wrap=lambda a: a, to_np=np.asarray, identity_fn=None, colors=None):
    """impl: object with splat_features / splat_features_from_scores / pyramid_resize.
    wrap: ndarray -> impl's array type (e.g. torch cuda tensor); to_np: the inverse."""
    z = arrays()
    ins = {k: wrap(v) for k, v in case_inputs(case).items()}
    kw = dict(case["kwargs"])
    for k in ("score_size", "viz_size"):
        if isinstance(kw.get(k), list):
            kw[k] = tuple(kw[k])
    if case["func"] == "splat_features":
        if case.get("uses_viz_score_fn"):
            kw["viz_score_fn"] = identity_fn or (lambda s: s)
        out = impl.splat_features(**ins, **kw)
        flat = _flatten(out, to_np)
    elif case["func"] == "splat_features_from_scores":
        flat = {"ret": to_np(impl.splat_features_from_scores(ins["scores"], ins["features"], kw["size"],
                                                            channels_last=kw["channels_last"]))}
    else:
        out = impl.pyramid_resize(ins["img"], kw["cutoff"])
        flat = {str(k): to_np(v) for k, v in out.items()}
    for key in case["outs"]:
        want = z[f"{case['name']}/out/{key}"]
        if key == "feature_img_sub":
            s = case["subsample"]
            got = flat["feature_img"][..., ::s, ::s]
        elif key == "feature_img_moments":
            img = flat["feature_img"].astype(np.float64)
            got = np.array([img.sum(), (img * img).sum(), img.max(), img.min()])
            want = want.astype(np.float64)
        elif "@c" in key:
            base, stride = key.split("@c")
            got = flat[base][:, ::int(stride)]
        else:
            got = flat[key]
        yield key, got, want


def check_close(got, want, rtol, atol, what=""):
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bound = atol + rtol * np.abs(want.astype(np.float64))
    bad = err > bound
    assert not bad.any(), (f"{what}: {bad.sum()} / {bad.size} out of tolerance; max abs err {err.max():.3e} "
                           f"at {np.unravel_index(err.argmax(), err.shape)} (want {want.flat[err.argmax()]!r})")
