"""Run the UNMODIFIED reference (yl-jiang/YOLOSeries) on seeded synthetic heads, in its own process.

TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.  The reference is imported from
/root/reference when that exists (build container) or from the byte-for-byte staged copy baseline/_ref/ (GPU box;
baseline/stage_reference.py), through oracle/refharness.py (which only stubs the plotting imports of utils/__init__.py).
Everything that is timed or compared is the reference's own code: ``XEvaluator.do_inference`` +
``XEvaluator.numba_nms`` (== ``XEvaluator.__call__`` with ``use_tta: false``, trainer/eval_yolov5.py:30-42,181-209,
261-316, utils/nms.py:10-27, utils/bbox_tools.py:12-35).

    python baseline/ref_worker.py case --family yolov5 --img 640 --dist dense --batch 1 --seed 7 --device cuda --out x.npz
        mode A (--device cpu) / mode B (--device cuda: ATen CUDA decode, host numba NMS -- the reference's deployment path)
    python baseline/ref_worker.py throughput --family yolov5 --img 640 --dist dense --procs 16 --steps 3 --budget 150
        P single-threaded processes (the reference is single-threaded: numba njit without parallel, OMP_NUM_THREADS=1
        under DDP, utils/setup_env.py:30-41), one image per process and step; prints one JSON line
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FCOS_OVER = {"compute_metric_cls_threshold": 0.2, "compute_metric_iou_threshold": 0.35, "max_predictions_per_img": 100,
             "cls_threshold": 0.2, "iou_threshold": 0.35}


def _single_thread():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    import torch
    torch.set_num_threads(1)


def build_case(family, img, dist, batch, seed, device, num_class=80, hyp_over=None):
    """-> (evaluator of the unmodified reference, dummy input batch, hyp)."""
    import torch

    from oracle import refharness
    from yoloseries_b200 import synth
    _utils, trainer = refharness.import_reference()
    over = dict(FCOS_OVER) if family == "fcos" else {}
    over.update(hyp_over or {})
    hyp = refharness.reference_hyp((img, img), num_class=num_class, device=device, **over)
    heads = synth.make_heads(family, batch, img, img, num_class, dist, seed, "cpu")  # CPU generator: same bits everywhere
    if device != "cpu":
        heads = _to_device(heads, device)
    ev = refharness.make_evaluator(trainer, family, refharness.head_model(family, heads), hyp)
    return ev, torch.zeros(batch, 3, img, img, device=device), hyp


def _to_device(x, device):
    import torch
    if isinstance(x, torch.Tensor):
        return x.to(device)
    return type(x)(_to_device(v, device) for v in x)


def pack_outputs(outs):
    """list[ndarray(K,6) | Tensor | None] -> (rows (b, Kmax, 6) f32, counts (b,) i32 with -1 == None)."""
    import numpy as np
    outs = [None if o is None else (o.cpu().numpy() if hasattr(o, "cpu") else np.asarray(o)) for o in outs]
    kmax = max([o.shape[0] for o in outs if o is not None] + [1])
    rows = np.zeros((len(outs), kmax, 6), dtype=np.float32)
    cnt = np.zeros(len(outs), dtype=np.int32)
    for i, o in enumerate(outs):
        if o is None:
            cnt[i] = -1
        else:
            cnt[i] = o.shape[0]
            rows[i, : o.shape[0]] = o.reshape(-1, 6)
    return rows, cnt


def run_cases(args):
    """--spec: JSON list of {name, family, img, dist, batch, seed, num_class, hyp}; one npz per case in --out (a dir).
    One process for all of them: the numba JIT (~12 s) is paid once."""
    import types
    with open(args.spec) as f:
        spec = json.load(f)
    os.makedirs(args.out, exist_ok=True)
    for c in spec:
        a = types.SimpleNamespace(family=c["family"], img=c["img"], dist=c["dist"], batch=c["batch"], seed=c["seed"],
                                  device=args.device, num_class=c.get("num_class", 80), hyp=json.dumps(c.get("hyp") or {}),
                                  out=os.path.join(args.out, c["name"] + ".npz"))
        run_case(a)


def run_case(args):
    import numpy as np
    import torch
    ev, dummy, hyp = build_case(args.family, args.img, args.dist, args.batch, args.seed, args.device, args.num_class,
                                json.loads(args.hyp) if args.hyp else None)
    t0 = time.perf_counter()
    decoded = ev.do_inference(dummy)
    if args.device != "cpu":
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    outs = ev.numba_nms(decoded.clone())
    t2 = time.perf_counter()
    rows, cnt = pack_outputs(outs)
    np.savez_compressed(args.out, decoded=decoded.float().cpu().numpy(), rows=rows, counts=cnt,
                        seconds=np.array([t1 - t0, t2 - t1]),
                        meta=np.array(repr({k: v for k, v in hyp.items() if isinstance(v, (int, float, bool, str))})))
    print(json.dumps({"decode_s": t1 - t0, "nms_s": t2 - t1, "counts": cnt.tolist()}))


# ---- throughput: P single-threaded worker processes ----------------------------------------------------------------
_W = {}


def _worker_init(family, img, dist, num_class):
    _single_thread()
    import torch
    # numba JIT + ATen warm-up on a small picture of the same family (compilation is per signature, not per size)
    ev, dummy, _ = build_case(family, 64 if family != "fcos" else 128, "dense", 1, 1, "cpu", num_class)
    ev(dummy)
    _W["torch"] = torch
    _W["cfg"] = (family, img, dist, num_class)
    _W["cases"] = {}


def _worker_loop(conn, family, img, dist, num_class, seed):
    """One single-threaded worker: import + JIT, build its image, then one reference call per 'go'."""
    os.environ["CUDA_VISIBLE_DEVICES"] = ""   # the CPU path: the reference's GPUAnchor would otherwise pick 'cuda'
    _worker_init(family, img, dist, num_class)
    ev, dummy, _ = build_case(family, img, dist, 1, seed, "cpu", num_class)   # inputs built outside the timing
    conn.send(("ready", os.getpid()))
    while True:
        msg = conn.recv()
        if msg != "go":
            break
        t0 = time.perf_counter()
        out = ev(dummy)   # XEvaluator.__call__: model stand-in -> do_inference -> numba_nms -> list[Tensor | None]
        dt = time.perf_counter() - t0
        conn.send((dt, -1 if out[0] is None else int(out[0].shape[0])))


def throughput(family, img, dist, procs, steps, warmup, budget_s, num_class=80, seed0=4321):
    """images/s of the unmodified reference with ``procs`` single-threaded processes, one image per process and step
    (the step ends when the slowest process is done).  Returns a dict."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t_setup = time.perf_counter()
    workers = []
    for i in range(procs):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_worker_loop, args=(b, family, img, dist, num_class, seed0 + i), daemon=True)
        pr.start()
        workers.append((pr, a))
    for _, a in workers:
        assert a.recv()[0] == "ready"
    setup_s = time.perf_counter() - t_setup      # untimed: process start, imports, numba JIT, input generation
    done, total, per_image, kept = 0, 0.0, [], []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        for _, a in workers:
            a.send("go")
        res = [a.recv() for _, a in workers]
        dt = time.perf_counter() - t0
        if k < warmup:
            continue
        done += 1
        total += dt
        per_image += [r[0] for r in res]
        kept = [r[1] for r in res]
        if total + dt > budget_s:
            break
    for pr, a in workers:
        a.send("stop")
    for pr, _ in workers:
        pr.join(timeout=10)
    return {"images_per_s": procs * done / total, "steps": done, "seconds": total, "procs": procs,
            "s_per_image_1core": sum(per_image) / len(per_image), "setup_s": setup_s, "kept_rows": kept}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["case", "cases", "throughput"])
    ap.add_argument("--spec", default="")
    ap.add_argument("--family", default="yolov5")
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--dist", default="dense")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--seed", type=int, default=4321)
    ap.add_argument("--num-class", type=int, default=80)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--hyp", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--budget", type=float, default=150.0)
    args = ap.parse_args()
    if args.mode == "case":
        run_case(args)
    elif args.mode == "cases":
        run_cases(args)
    else:
        print(json.dumps(throughput(args.family, args.img, args.dist, args.procs, args.steps, args.warmup, args.budget,
                                    args.num_class, args.seed)))


if __name__ == "__main__":
    main()
