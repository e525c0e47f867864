#!/usr/bin/env python3
"""bench.py — sentences/sec of top-1 analysis at batch 8192 (BASELINE.json config[1]) on N B200s.

One "step" = one pass of the hot path (lattice kernel + Viterbi kernel + pack) over one batch of 8192 synthetic
Korean sentences with the web.txt length distribution (kiwi_b200/synth.py), fabricated Knlm model.
  value  : inputs already resident in HBM, device-timed (CUDA events on the engine stream), whole job.
  e2e    : the same batch through the public C-ABI call kiwi_b200_analyze_batch with HOST buffers
           (pinned staging, H2D of text+offsets and D2H of the packed tokens inside the timed region).
  --impl reference : the reference's own CPU implementation (oracle/_ref, all host threads) on a bounded sample.
Launch: python bench.py --gpus N --steps K --warmup W   (N > 1: under torchrun, one rank per GPU)."""
import argparse, json, os, subprocess, sys, tempfile, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
IMAGE = os.path.join(ROOT, "oracle", "_ref", "models", "knlm_small.img")
REF_MODEL_DIR = os.path.join(ROOT, "oracle", "_ref", "models", "knlm_small")
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench")
METRIC = "sentences/sec (top-1 analyze, batch=8192)"
UNIT = "sentences/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe, clocks line)."""
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.stop_flag = False; self.samples = []; self.reasons = set(); self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"): self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def set_model(name):
    """knlm (headline, BASELINE.json config[1]) or cong (config[2]: quantized CoNg model, int8 scorer on tensor-core tiles)"""
    global IMAGE, REF_MODEL_DIR, MODEL
    MODEL = name
    IMAGE = os.path.join(ROOT, "oracle", "_ref", "models", name + "_small.img")
    REF_MODEL_DIR = os.path.join(ROOT, "oracle", "_ref", "models", name + "_small")


MODEL = "knlm"


def work_counters(batch_texts, batch_size, seed):
    """Per-sentence algorithmic bytes from the instrumented oracle (committed under profiles/; recomputed on a sample if absent)."""
    from kiwi_b200 import bytemodel
    name = "counters_r1.json" if MODEL == "knlm" else "counters_r1b_cong.json"
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        d = json.load(open(path))
        key = "batch%d_seed%d" % (batch_size, seed)
        if key in d:
            c = d[key]
            extra = bytemodel.cong_bytes(c) if MODEL == "cong" else 0.0
            return c, bytemodel.lattice_bytes(c) / c["sentences"], (bytemodel.viterbi_bytes(c) + extra) / c["sentences"], "profiles/" + name
    from tests.orc import Oracle
    o = Oracle(IMAGE)
    sample = batch_texts[:256]
    for s in sample: o.analyze(s)
    c = o.work_counters()
    if MODEL == "cong": c.update(o.cong_counters())
    o.close()
    extra = bytemodel.cong_bytes(c) if MODEL == "cong" else 0.0
    return c, bytemodel.lattice_bytes(c) / c["sentences"], (bytemodel.viterbi_bytes(c) + extra) / c["sentences"], "oracle sample of 256 sentences"


def run_reference_cpu(texts, threads, repeats=1):
    """Times the unmodified reference (oracle/_ref) when it travelled with the repo, else the oracle port."""
    if os.path.exists(REF_BENCH) and os.path.exists(REF_MODEL_DIR):
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False, encoding="utf-8") as f:
            for t in texts: f.write(t + "\n")
            tmp = f.name
        try:
            env = dict(os.environ); env.setdefault("KIWI_ARCH_TYPE", "avx2"); env["KB_MODEL_TYPE"] = MODEL
            out = subprocess.run([REF_BENCH, REF_MODEL_DIR, tmp, str(threads), str(repeats)], capture_output=True, text=True, timeout=1200, env=env)
            line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
            r = json.loads(line)
            return {"value": r["sent_per_s"], "unit": UNIT, "cores": threads, "kind": "reference",
                    "sample": "%d sentences of the bench batch, reference Kiwi::analyze batch mode on %d threads, KIWI_ARCH_TYPE=%s, best of %d passes: %.2f s" % (r["sentences"], threads, r["arch"], r["repeats"], r["seconds"])}
        finally:
            os.unlink(tmp)
    from tests.orc import Oracle
    o = Oracle(IMAGE)
    t0 = time.time()
    for t in texts: o.analyze(t)
    dt = time.time() - t0
    return {"value": len(texts) / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": "%d sentences, oracle restatement single thread, %.2f s" % (len(texts), dt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (kernel experiments)")
    ap.add_argument("--model", default="knlm", choices=["knlm", "cong"], help="knlm = headline config; cong = BASELINE.json config[2] (CoNg model)")
    args = ap.parse_args()
    set_model(args.model)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from kiwi_b200.synth import synth_batch, SEED

    if args.impl == "reference":
        if rank != 0: return 0
        threads = os.cpu_count() or 1
        texts = synth_batch(args.batch, SEED)[:args.cpu_sample]
        vals = []
        for _ in range(max(1, args.warmup > 0)): run_reference_cpu(texts[:256], threads)
        last = None
        t0 = time.time()
        for _ in range(args.steps):
            last = run_reference_cpu(texts, threads); vals.append(last["value"])
            if time.time() - t0 > 240: break
        v = sum(vals) / len(vals)
        last["value"] = v
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
                          "ms_per_step": 1000.0 * len(texts) / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "batch=%d synthetic Korean sentences (web.txt length dist), fabricated %s model, top-1; each step = %d-sentence sample on %d host threads" % (args.batch, MODEL, len(texts), threads)},
                          "cpu_baseline": last, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import numpy as np
    import torch
    import kiwi_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: kiwi_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # ---- model: rank 0 reads the image, one NCCL broadcast puts it on every GPU (no collectives afterwards)
    if rank == 0:
        img = np.fromfile(IMAGE, dtype=np.uint8)
        size = torch.tensor([img.size], dtype=torch.int64, device="cuda")
    else:
        size = torch.zeros(1, dtype=torch.int64, device="cuda")
    if world > 1: dist.broadcast(size, 0)
    buf = torch.from_numpy(img).cuda() if rank == 0 else torch.empty(int(size.item()), dtype=torch.uint8, device="cuda")
    if world > 1: dist.broadcast(buf, 0)
    image_bytes = buf.cpu().numpy().tobytes()
    del buf
    kw = kiwi_b200.Kiwi(image_bytes=image_bytes, device=local_rank)

    # ---- per-rank synthetic batches (weak scaling: every rank works on `batch` sentences per step)
    R = 4
    batches = []
    for r in range(R):
        texts = synth_batch(args.batch, SEED + 1000 * rank + r)
        blob, off = kiwi_b200.encode_batch(texts)
        d_blob = torch.from_numpy(blob.view(np.int16)).cuda(); d_off = torch.from_numpy(off.view(np.int32)).cuda()
        batches.append((texts, blob, off, d_blob, d_off))
    torch.cuda.synchronize()

    def step_device(i):
        _, blob, off, d_blob, d_off = batches[i % R]
        return kw.analyze_device(d_blob.data_ptr(), d_off.data_ptr(), args.batch, int(blob.size))

    for i in range(args.warmup): step_device(i)
    sampler = ClockSampler(local_rank); sampler.start()
    if dist: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    dev_ms = 0.0; ms_vit = 0.0; ms_lat = 0.0; launches = 0; tokens = 0; retried = 0
    for i in range(args.steps):
        ms, nt, nl = step_device(i)
        st = kw.last_stats()
        dev_ms += ms; ms_vit += st.ms_viterbi; ms_lat += st.ms_lattice; launches += st.kernel_launches; tokens += nt; retried += st.retried
    torch.cuda.synchronize()
    if dist: dist.barrier()
    wall = time.time() - t0
    sampler.stop_flag = True; sampler.join(timeout=2)

    # ---- end to end through the public API with host buffers
    for i in range(min(2, args.warmup)): kw.analyze_batch_arrays(batches[i % R][1], batches[i % R][2])
    if dist: dist.barrier()
    torch.cuda.synchronize()
    t1 = time.time(); h2d = d2h = 0
    for i in range(args.steps):
        kw.analyze_batch_arrays(batches[i % R][1], batches[i % R][2])
        st = kw.last_stats(); h2d += st.h2d_bytes; d2h += st.d2h_bytes
    torch.cuda.synchronize()
    if dist: dist.barrier()
    e2e_wall = time.time() - t1

    t = torch.tensor([dev_ms, wall, e2e_wall, ms_vit, ms_lat], dtype=torch.float64, device="cuda")
    if dist: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall, e2e_wall, ms_vit, ms_lat = [float(x) for x in t.tolist()]
    if rank != 0:
        if dist: dist.destroy_process_group()
        return 0

    total_sent = args.batch * args.steps * world
    value = total_sent / (dev_ms / 1000.0)
    e2e = total_sent / e2e_wall
    peak, peak_kind = load_peaks()
    c, lat_b, vit_b, csrc = work_counters(batches[0][0], args.batch, SEED)
    vit_per_launch_ms = ms_vit / args.steps
    achieved = vit_b * args.batch / (vit_per_launch_ms / 1000.0) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic_r1b.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("viterbi_kernel_dram_bytes_per_launch" if MODEL == "knlm" else "viterbi_cong_kernel_dram_bytes_per_launch")
    cpu = run_reference_cpu(batches[0][0][:args.cpu_sample], os.cpu_count() or 1, repeats=5) if (world == 1 and not args.no_cpu) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch=%d synthetic Korean sentences (web.txt length dist) per GPU per step, fabricated %s model (%s_small), top-1" % (args.batch, "Knlm" if MODEL == "knlm" else "CoNg", MODEL),
                   "l2": "per-step scratch working set (GBs) exceeds the 126 MB L2 and %d distinct input batches rotate; the read-only model stays resident as in steady state" % R,
                   "parallelism": "dp%d (sentence sharding, one NCCL broadcast of the model image at init, no steady-state collectives)" % world,
                   "wall_ms_per_step": 1000.0 * wall / args.steps,
                   "retried_sentences_per_step": retried / args.steps},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": "viterbi_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6650",
                     "algorithmic_bytes_per_sentence": vit_b, "lattice_kernel_bytes_per_sentence": lat_b, "counters": csrc,
                     "kernel_ms_per_launch": vit_per_launch_ms, "lattice_ms_per_launch": ms_lat / args.steps,
                     "kernel_share_of_step": ms_vit / dev_ms if dev_ms else None},
        "tokens_per_step": tokens // args.steps,
    }
    if cpu: line["cpu_baseline"] = cpu
    if MODEL == "cong":
        # the int8 gather GEMMs of progressMatrix: 2 * sum(m * n * dim) integer ops per sentence (counted by the instrumented oracle)
        ops = 2.0 * c.get("cgMacs", 0) / max(1, c["sentences"])
        line["roofline"]["kernel"] = "viterbi_cong_kernel"
        line["roofline"]["tensor"] = {"bound": "tensor", "unit": "TOP/s", "achieved": ops * args.batch / (vit_per_launch_ms / 1000.0) / 1e12, "peak": 4500.0,
                                      "peak_kind": "nominal dense int8 (no measured int8 peak in MEASURED_PEAKS.json)", "int8_ops_per_sentence": ops,
                                      "note": "the gather GEMMs are a small part of viterbi_cong_kernel's time; the kernel as a whole is latency / HBM-L2 bound like the Knlm one"}
    print(json.dumps(line))
    if dist: dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
