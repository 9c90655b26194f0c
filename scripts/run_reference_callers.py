#!/usr/bin/env python
"""Run the reference's own callers -- ``get_basis.py`` and ``train.py``, UNMODIFIED, as ``__main__`` with their own
argparse, data loaders and loops -- over a synthetic S3DIS-format data set on disk, against

  --impl dropin      this repo's ``model/``, ``runs/eval.py`` and ``gfs3d.kmeans.KMeans`` (put in place of
                     ``sklearn.cluster.KMeans``), i.e. the claim "the callers run unchanged";
  --impl reference   the reference's own ``model/``, ``runs/eval.py`` and scikit-learn, stock PyTorch on the same GPU.

TEST INFRASTRUCTURE: the reference sources are the byte copies staged by ``oracle/make_ref.py`` under
``oracle/_ref/reference`` (git-ignored).  Nothing here is imported by the product.

Environment drift that is shimmed, none of it on the hot path: ``h5py`` / ``transforms3d`` are imported by
``dataloaders/loader.py:10-11`` but not installed (h5py is never used; transforms3d only under ``--pc_augm``) -> empty stub
modules; ``np.int`` (``loader.py:104``, removed in numpy 1.24) -> ``int``; ``torch.load`` defaults to ``weights_only=True``
since torch 2.6 while ``train.py`` stores a numpy scalar in its checkpoint -> TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1.

Usage
  python scripts/run_reference_callers.py --work /tmp/gfs_callers            # whole pipeline, prints a transcript
  (child mode, used internally)  --child --impl dropin --script train.py -- <the script's own argv>
"""
import argparse
import json
import os
import pickle
import re
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gfs-3dseg_gws_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "reference")

S3DIS_NAMES = ["ceiling", "floor", "wall", "beam", "column", "window", "door", "table", "chair", "sofa", "bookcase",
               "board", "clutter"]


# ------------------------------------------------------------------------------------------------------------------
# child: one reference script as __main__
# ------------------------------------------------------------------------------------------------------------------
def child(impl: str, script: str, argv):
    import runpy

    import numpy as np
    if not hasattr(np, "int"):
        np.int = int                                         # dataloaders/loader.py:104
    for name in ("h5py", "transforms3d"):                    # dataloaders/loader.py:10-11
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
    import torch
    torch.backends.cudnn.allow_tf32 = False                  # the reference arm computes in true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    if os.environ.get("GFS_CALLERS_CPU_DEBUG") == "1":       # build-container dry run of the reference arm only (no GPU here)
        assert impl == "reference"
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.empty_cache = lambda: None
        torch.cuda.manual_seed_all = lambda s: None

    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, os.path.join(ROOT, "scripts"))]
    if impl == "dropin":
        sys.path[:0] = [PKG, REF]                            # model.*, runs.eval -> this repo; util, dataloaders -> reference
        import sklearn.cluster
        from gfs3d.kmeans import KMeans
        sklearn.cluster.KMeans = KMeans                      # get_basis.py:19 `from sklearn.cluster import KMeans`
    else:
        sys.path[:0] = [REF]
    sys.argv = [os.path.join(REF, script)] + list(argv)
    runpy.run_path(os.path.join(REF, script), run_name="__main__")
    import model.dgcnn as md
    import runs.eval as re_
    print(f"[callers] impl={impl} model.dgcnn={md.__file__} runs.eval={re_.__file__}")


# ------------------------------------------------------------------------------------------------------------------
# synthetic S3DIS-format data set on disk (what dataloaders/s3dis.py:51-84 and loader.py:39-129 read)
# ------------------------------------------------------------------------------------------------------------------
def make_dataset(root: str, n_blocks: int, seed: int, pts=(2300, 3000)):
    """root/blocks/data/<name>.npy rows = [x y z r g b label]; root/meta/s3dis_classnames.txt.  Every class is a
    recognisable thing (own height band, colour, footprint) so a few epochs of training give a non-trivial mIoU."""
    import numpy as np
    rng = np.random.RandomState(seed)
    data_dir = os.path.join(root, "blocks", "data")
    os.makedirs(data_dir, exist_ok=True)
    os.makedirs(os.path.join(root, "meta"), exist_ok=True)
    with open(os.path.join(root, "meta", "s3dis_classnames.txt"), "w") as f:
        f.write("\n".join(S3DIS_NAMES) + "\n")
    look = np.random.RandomState(7)                              # what a class looks like: the same in every data set
    colour = look.uniform(30, 225, size=(13, 3))
    zlo = np.linspace(0.0, 2.4, 13)[look.permutation(13)]
    for b in range(n_blocks):
        n = int(rng.randint(*pts))
        classes = [b % 13, (5 * b + 3) % 13, (7 * b + 6) % 13, int(rng.randint(13))]
        classes = list(dict.fromkeys(classes))
        share = rng.dirichlet(np.full(len(classes), 4.0))
        lab = np.repeat(classes, np.maximum((share * n).astype(int), 160))[:n]
        n = lab.shape[0]
        cx = rng.uniform(0.15, 0.85, size=(13, 2))
        xy = cx[lab] + rng.normal(0, 0.08, size=(n, 2)) * (1 + (lab % 3))[:, None]
        z = zlo[lab] + rng.uniform(0, 0.5, size=n) * (1 + (lab % 2))
        rgb = np.clip(colour[lab] + rng.normal(0, 12, size=(n, 3)), 0, 255)
        arr = np.concatenate([np.clip(xy, 0, 1), z[:, None], rgb, lab[:, None].astype(np.float64)], axis=1)
        np.save(os.path.join(data_dir, f"Area_{1 + b % 5}_room_{b:03d}_block_{b}.npy"), arr[rng.permutation(n)])
    return os.path.join(root, "blocks")


def make_checkpoint(path: str):
    """<path>/checkpoint.tar = {'params': encoder state dict}  (util/checkpoint_util.py:10-17,52-53)"""
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", "gfs_s3dis_weights.npz"))
    enc = {k[len("encoder."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("encoder.")}
    os.makedirs(path, exist_ok=True)
    torch.save(dict(params=enc), os.path.join(path, "checkpoint.tar"))


def run_child(impl, script, argv, log):
    if os.environ.get("GFS_CALLERS_CPU_DEBUG") == "1":
        impl = "reference"
    cmd = [sys.executable, os.path.abspath(__file__), "--child", "--impl", impl, "--script", script, "--"] + argv
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    dt = time.time() - t0
    with open(log, "w") as f:
        f.write(p.stdout)
    return p.returncode, p.stdout, dt


def pipeline(work: str, epochs: int = 5, num_cnt: int = 150, say=print):
    """get_basis.py (both impls) -> train.py --epochs (drop-in) -> train.py --only_evaluate (both impls, same checkpoint)"""
    import numpy as np
    os.makedirs(work, exist_ok=True)
    res = {}
    train_root = make_dataset(os.path.join(work, "S3DIS_train"), 78, seed=11)
    test_root = make_dataset(os.path.join(work, "S3DIS_test"), 26, seed=12)
    make_checkpoint(os.path.join(work, "pretrain"))
    common = ["--dataset", "s3dis", "--cvfold", "0", "--data_path", train_root, "--n_workers", "0",
              "--pretrain_checkpoint_path", os.path.join(work, "pretrain")]

    # ---- get_basis.py as __main__ (get_basis.py:226-311 -> Get_GlobalProto_GlobalKmeans :112-222)
    for impl in ("dropin", "reference"):
        sp = os.path.join(work, f"basis_{impl}")
        rc, out, dt = run_child(impl, "get_basis.py", common + ["--num_cnt", str(num_cnt), "--save_path", sp, "--seed", "123"],
                                os.path.join(work, f"log_get_basis_{impl}.txt"))
        assert rc == 0, out[-3000:]
        fn = os.path.join(sp, f"GlobalKmeans_EdgeConv123_cnt={num_cnt}_energy=095_SVDReconstruct.pkl")
        basis = pickle.load(open(fn, "rb"))
        km = re.search(r"kmean : ([0-9.eE+-]+)", out)
        res[f"basis_{impl}"] = {"shape": list(basis.shape), "dtype": str(basis.dtype), "finite": bool(np.isfinite(basis).all()),
                                "kmeans_s": float(km.group(1)) if km else None, "wall_s": round(dt, 1), "file": fn}
        say(f"get_basis.py [{impl}]: basis {basis.shape} {basis.dtype}, k-means {res[f'basis_{impl}']['kmeans_s']} s, {dt:.1f} s wall")
    A = pickle.load(open(res["basis_dropin"]["file"], "rb")).astype(np.float64)
    Bm = pickle.load(open(res["basis_reference"]["file"], "rb")).astype(np.float64)
    An, Bn = A / np.linalg.norm(A, axis=1, keepdims=True), Bm / np.linalg.norm(Bm, axis=1, keepdims=True)
    cs = An @ Bn.T
    res["basis_nearest_word_cosine"] = {"mean": float(cs.max(1).mean()), "min": float(cs.max(1).min())}
    res["basis_rank"] = {"dropin": int(np.linalg.matrix_rank(A, tol=1e-4 * np.linalg.norm(A, 2))),
                         "reference": int(np.linalg.matrix_rank(Bm, tol=1e-4 * np.linalg.norm(Bm, 2)))}
    say(f"  geometric words: every drop-in word's nearest reference word has cosine mean {cs.max(1).mean():.4f} / min {cs.max(1).min():.4f}; "
        f"rank after the 95 % energy cut {res['basis_rank']}")

    # ---- train.py as __main__: coding collection, `epochs` training epochs, support prototypes, validation, checkpoint
    save = os.path.join(work, "run_dropin")
    targv = common + ["--testing_data_path", test_root, "--save_path", save, "--use_pretrain_weight", "--batch_size", "8",
                      "--epochs", str(epochs), "--basis_path", res["basis_dropin"]["file"], "--energy", "0.9", "--print_freq", "1",
                      "--total_classes", "13", "--k_shot", "5", "--base_lr", "0.01"]
    rc, out, dt = run_child("dropin", "train.py", targv, os.path.join(work, "log_train_dropin.txt"))
    assert rc == 0, out[-3000:]
    losses = [float(m) for m in re.findall(r"Loss ([0-9.]+) \(", out)]
    accs = [float(m) for m in re.findall(r"Train result at epoch \[\d+/\d+\]: acc (\d+\.\d+)", out)]
    ev = re.search(r"Epoch: (\d+), Final mIoU: ([0-9.eE+-]+), BASE: ([0-9.eE+-]+), NOVEL: ([0-9.eE+-]+), hm: ([0-9.eE+-]+)", out)
    ckpts = sorted(f for f in os.listdir(save) if f.startswith("train_epoch_") and f.endswith(".pth"))
    res["train_dropin"] = {"iterations": len(losses), "loss_first": losses[0], "loss_last": losses[-1], "epoch_acc": accs,
                           "val": [float(ev.group(i)) for i in range(2, 6)] if ev else None, "checkpoints": ckpts, "wall_s": round(dt, 1)}
    say(f"train.py [dropin]: {len(losses)} iterations over {epochs} epochs, loss {losses[0]:.4f} -> {losses[-1]:.4f}, "
        f"epoch accuracy {accs}, validation (mIoU, base, novel, hm) {res['train_dropin']['val']}, saved {ckpts}, {dt:.1f} s wall")
    assert ckpts, "train.py saved no checkpoint (validation mIoU was 0)"

    # ---- train.py --only_evaluate as __main__ on THAT checkpoint, with both implementations
    for impl in ("dropin", "reference"):
        eargv = [a for a in targv if a != "--use_pretrain_weight"] + [
            "--only_evaluate", "--model_checkpoint_path", os.path.join(save, ckpts[-1]), "--eval_weight", "1.2", "--phase", "test"]
        rc, out, dt = run_child(impl, "train.py", eargv, os.path.join(work, f"log_eval_{impl}.txt"))
        assert rc == 0, out[-3000:]
        m = re.search(r"Eval result: Final mIoU: ([0-9.eE+-]+), BASE: ([0-9.eE+-]+), NOVEL: ([0-9.eE+-]+), hm_mIoU: ([0-9.eE+-]+)", out)
        cls = [float(x) for x in re.findall(r"class \d+, iou over multiple runs: ([0-9.eE+-]+)", out)]
        res[f"eval_{impl}"] = {"mIoU": float(m.group(1)), "base": float(m.group(2)), "novel": float(m.group(3)), "hm": float(m.group(4)),
                               "class_iou": cls, "wall_s": round(dt, 1)}
        say(f"train.py --only_evaluate [{impl}] (5 support seeds): mIoU {m.group(1)}, base {m.group(2)}, novel {m.group(3)}, hm {m.group(4)}, {dt:.1f} s wall")
    d, r = res["eval_dropin"], res["eval_reference"]
    res["eval_abs_diff"] = {k: abs(d[k] - r[k]) for k in ("mIoU", "base", "novel", "hm")}
    res["eval_abs_diff"]["class_iou_max"] = max(abs(a - b) for a, b in zip(d["class_iou"], r["class_iou"]))
    say(f"  |drop-in - reference| on the same checkpoint: {res['eval_abs_diff']}")
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--impl", default="dropin")
    ap.add_argument("--script", default="train.py")
    ap.add_argument("--work", default="/tmp/gfs_callers")
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--json", default=None)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    if a.child:
        child(a.impl, a.script, a.rest[1:] if a.rest[:1] == ["--"] else a.rest)
    else:
        if not os.path.exists(os.path.join(REF, "MANIFEST.json")):
            sys.exit("oracle/_ref/reference is not staged (run python oracle/make_ref.py where /root/reference exists)")
        out = pipeline(a.work, a.epochs)
        if a.json:
            json.dump(out, open(a.json, "w"), indent=1)
