#!/usr/bin/env python3
"""bench.py — decoded megapixels/sec of the B200 JPEG path (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload restart8|norestart|4k444rgb|encode]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the reference's CPU algorithm (oracle port) on the host cores

A step = one pass of the hot path over one batch of synthetic images:
  value   whole-job MP/s with the compressed batch already resident in HBM (kernels + the
          coefficient-buffer clear only), CUDA events on the library's stream, max over ranks;
  e2e     the same metric through hcj_decode_batch with pinned HOST buffers: header parse, H2D of the
          files, kernels, D2H of the frames, all inside the timed region (wall clock around the call);
  roofline  algorithmic bytes of the dominant kernel / its CUDA-event duration vs MEASURED_PEAKS.json;
  cpu_baseline  the CPU oracle (a port of the OCaml model) on the host cores, one process per core,
          on a bounded sample of the same images.
Images are sharded by batch index (weak scaling: `--batch` images per GPU), no collective on the data path.
The synthetic JPEGs are produced by the product's own GPU encoder (byte-identical to the model's encoder,
tests/test_gpu_encode.py); oracle/ is executed only for cpu_baseline / --impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "video-coding_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (width, height, chroma, quality, restart_interval, out mode name, default batch per GPU)
    "restart8": (1920, 1080, 420, 75, 8, "yuv", 1024),   # BASELINE configs[1]
    "norestart": (1920, 1080, 420, 75, 0, "yuv", 1024),  # BASELINE configs[2] (8192 / 8 GPUs)
    "4k444rgb": (3840, 2160, 444, 95, 0, "rgb", 128),    # BASELINE configs[3]
    "restart8rgb": (1920, 1080, 420, 75, 8, "rgb", 1024),  # configs[1] with RGB24 output (Planar_444 up-sampling + colour)
    "encode": (1920, 1080, 420, 75, 0, "jpeg", 512),     # BASELINE configs[4]
}


def _synth_one(args):
    import synth

    seed, w, h, chroma = args
    return synth.frame(seed, w, h, chroma)


def make_frames(n, w, h, chroma, seed0):
    """`n` distinct seeded frames (SURVEY 8d recipe), generated on the host cores before CUDA starts."""
    import multiprocessing as mp

    jobs = [(seed0 + i, w, h, chroma) for i in range(n)]
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        return pool.map(_synth_one, jobs)


def _cpu_decode_worker(args):
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as orc

    jpgs, reps = args
    orc.lib()
    t = orc.time_decode(jpgs, True, reps)
    return t


def _cpu_encode_worker(args):
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as orc

    frames, w, h, chroma, q, ri = args
    orc.lib()
    return sum(orc.time_encode(f, w, h, chroma, q, ri, 1) for f in frames)


def cpu_arm(kind, items, w, h, chroma, q, ri, target_s=12.0):
    """Time the oracle (port of the model's algorithm) with one process per host core on a bounded sample."""
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    per_image = 0.07 * (w * h) / (1920 * 1080) * (3.0 if chroma == 444 else 1.0)
    per_core = max(1, min(64, int(target_s / per_image / 1.0)))
    ctx = mp.get_context("spawn")
    chunks = [[items[(c * per_core + i) % len(items)] for i in range(per_core)] for c in range(cores)]
    with ctx.Pool(cores) as pool:
        pool.map(_noop, range(cores))  # start the workers (imports) outside the timed region
        t0 = time.perf_counter()
        if kind == "decode":
            pool.map(_cpu_decode_worker, [(ch, 1) for ch in chunks])
        else:
            pool.map(_cpu_encode_worker, [(ch, w, h, chroma, q, ri) for ch in chunks])
        wall = time.perf_counter() - t0
    n = per_core * cores
    return {
        "value": n * w * h / 1e6 / wall,
        "unit": "MP/s",
        "cores": cores,
        "kind": "port",
        "sample": "%d images (%d per core, one process per core, oracle/hcj_oracle.c -O2), %.1f s wall" % (n, per_core, wall),
    }


def _noop(_):
    return 0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index, self.first = [], None, index, 0

    def mark(self):
        """Samples before this point (input generation, warm-up of nvidia-smi itself) are ignored."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.rows = self.rows[self.first:]
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic(kernel, n_images):
    """dram__bytes_read + dram__bytes_write of `kernel` for one launch over n_images images, from the committed
    ncu --set full capture (profiles/r01s3_traffic.json: per-image bytes on the same synthetic 1080p images; the
    capture ran 296 images per launch).  None for kernels / workloads that were not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01s3_traffic.json")) as f:
            per_image = json.load(f)["dram_bytes_per_image"].get(kernel)
        return None if per_image is None else per_image * n_images
    except Exception:
        return None


def bind_to_gpu_numa_node(index):
    """Pin this process (and so its pinned host buffers, first-touch) to the CPUs of the NUMA node the GPU hangs
    off: with one process per GPU the PCIe copies of the end-to-end leg then stay on the GPU's own socket.
    Best effort: returns the node or None."""
    try:
        bdf = subprocess.check_output(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      text=True).strip().lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


_REAL_STDOUT = None


def emit(obj):
    """The one JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to stdout on their own (NCCL announces its version there at communicator creation): everything
    # but the JSON line goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="restart8", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: per workload)")
    ap.add_argument("--unique", type=int, default=64, help="distinct images cycled to fill the batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w, h, chroma, quality, ri, out_name, default_batch = WORKLOADS[args.workload]
    batch_n = args.batch or default_batch
    unique = max(1, min(args.unique, batch_n))
    config = {"workload": "%s: %d x %dx%d %d q%d %s per GPU -> %s" % (
        args.workload, batch_n, w, h, chroma, quality, "DRI=%d MCUs" % ri if ri else "no restart markers", out_name),
        "batch_per_gpu": batch_n, "unique_images": unique, "sharding": "batch index, no collective",
        "l2": "per-step working set (compressed batch + coefficient buffer) exceeds the 126 MB L2"}

    if args.impl == "reference" and rank != 0:
        return 0

    # ---- synthetic inputs (host cores, before CUDA is touched)
    frames = make_frames(unique if args.impl == "ours" else min(unique, 16), w, h, chroma, 1000 * (1 + rank))

    import hcjpeg

    hcjpeg.build()
    if args.impl == "reference":
        # The reference's own CPU implementation cannot run here (OCaml, no toolchain): the oracle port of its
        # algorithm is timed on all host cores.  Inputs are produced by the GPU encoder when a GPU is present
        # (byte-identical), else by the oracle encoder.
        from oracle import pyoracle as orc

        kind = "encode" if args.workload == "encode" else "decode"
        if kind == "decode":
            items = [orc.encode(f, w, h, chroma, quality, restart_interval=ri) for f in frames]
        else:
            items = frames
        vals = []
        for step in range(args.warmup + args.steps):
            r = cpu_arm(kind, items, w, h, chroma, quality, ri, target_s=8.0)
            if step >= args.warmup:
                vals.append(r)
        v = float(np.mean([x["value"] for x in vals]))
        cb = dict(vals[-1], value=v)
        emit(({
            "impl": "reference", "metric": "%s megapixels/sec (%dx%d %d baseline)" % ("decoded" if kind == "decode" else "encoded", w, h, chroma),
            "value": v, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean([float(x["sample"].split(",")[-1].split()[0]) for x in vals])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": config, "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    config["numa_node"] = numa
    import torch

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from hcjpeg import shard

    dist_mod = dist if world > 1 else None

    def barrier():
        shard.barrier(dist_mod, torch.cuda.synchronize)

    def max_over_ranks(x):
        return shard.max_over_ranks(x, dist_mod, "cuda")

    ctx = hcjpeg.Context(local)
    mode = {"yuv": hcjpeg.OUT_YUV, "rgb": hcjpeg.OUT_RGB24, "jpeg": None}[out_name]
    mp_per_step = batch_n * w * h / 1e6
    sampler = ClockSampler(local)
    sampler.start()
    result = {}

    if args.workload != "encode":
        # compressed inputs from the GPU encoder (byte-identical to the model's encoder)
        jpgs, st = [], []
        for i in range(0, unique, 16):
            o, s = ctx.encode_batch(frames[i:i + 16], w, h, chroma, quality, ri)
            jpgs += o
            st += s
        assert all(s == 0 for s in st), st
        batch_jpgs = [jpgs[i % unique] for i in range(batch_n)]
        comp_bytes = sum(len(j) for j in batch_jpgs)

        # ---- value: resident inputs, kernels only
        b = ctx.batch(batch_jpgs, mode)
        assert all(s == 0 for s in b.host_status)
        sampler.mark()
        for _ in range(max(args.warmup, 3)):
            b.decode()
        ctx.synchronize()
        barrier()
        ctx.timer_start()
        for _ in range(args.steps):
            b.decode()
        ms = ctx.timer_stop()
        barrier()
        ms = max_over_ranks(ms)
        ms_per_step = ms / args.steps
        # per-stage timing for the roofline of the dominant kernel (separate passes, same stream, CUDA events)
        stages = {}
        for _ in range(args.steps):
            for k, v in b.decode_stages().items():
                stages[k] = stages.get(k, 0.0) + v / args.steps
        clocks = sampler.stop()
        outs, st = b.fetch()
        assert all(s == 0 for s in st), [s for s in st if s][:4]
        nblocks = sum(f.nblocks for f in b.infos)
        out_bytes = sum(hcjpeg.out_size(f, mode) for f in b.infos)
        launches = b.kernels() * args.steps
        b.close()
        alg = {  # algorithmic bytes per launch (SURVEY 8d / DESIGN.md)
            "destuff": 2 * comp_bytes,
            "huffman_restart": comp_bytes + 128 * nblocks,
            "huffman_speculative": comp_bytes + 128 * nblocks,
            "idct": 128 * nblocks + (out_bytes if mode != hcjpeg.OUT_RGB24 else nblocks * 64),
            "rgb": nblocks * 64 + out_bytes,
            "clear_flags": nblocks // 8,
        }
        stages = {k: v for k, v in stages.items() if v > 0.02}  # stages that did not launch read as ~0.003 ms
        kernels_only = {k: v for k, v in stages.items() if k != "clear_flags"}
        dom = max(kernels_only, key=kernels_only.get)
        peak, peak_src = peaks()
        achieved = alg[dom] / (stages[dom] * 1e-3) / 1e9
        traffic = measured_traffic(dom, batch_n) if (w, h, chroma, quality) == (1920, 1080, 420, 75) else None
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": alg[dom],
                    "traffic_source": "profiles/r01s3_traffic.json (ncu --set full, dram read + write per image x batch)" if traffic else None,
                    "peak_source": peak_src,
                    "stages": {k: {"ms": v, "algorithmic_GBps": alg[k] / (v * 1e-3) / 1e9 if v > 0 else None,
                                   "frac": alg[k] / (v * 1e-3) / 1e9 / peak if v > 0 else None} for k, v in stages.items()}}
        result.update(ms_per_step=ms_per_step, value=world * mp_per_step / (ms_per_step * 1e-3), roofline=roofline,
                      gpu_launches=launches, clocks=clocks)

        # ---- e2e: host buffers in and out through hcj_decode_batch
        if not args.no_e2e:
            import ctypes as C

            L = hcjpeg.lib()
            pin_in = L.hcj_host_alloc(comp_bytes + 64 * batch_n)
            pin_out = L.hcj_host_alloc(out_bytes + 256 * batch_n)
            assert pin_in and pin_out
            jp = (C.c_void_p * batch_n)()
            lens = (C.c_size_t * batch_n)()
            op = (C.c_void_p * batch_n)()
            caps = (C.c_size_t * batch_n)()
            status = (C.c_int * batch_n)()
            off_i = off_o = 0
            for i, j in enumerate(batch_jpgs):
                C.memmove(pin_in + off_i, j, len(j))
                jp[i], lens[i] = pin_in + off_i, len(j)
                off_i += (len(j) + 16 + 15) // 16 * 16
                size = hcjpeg.out_size(hcjpeg.frame_info(j), mode) if i < unique else caps[i % unique]
                op[i], caps[i] = pin_out + off_o, size
                off_o += (size + 255) // 256 * 256
            e2e_steps = max(1, min(args.steps, 3))
            for _ in range(1):
                hcjpeg._check(L.hcj_decode_batch(ctx._h, jp, lens, batch_n, mode, hcjpeg.FLAG_DEFAULT, op, caps, status))
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                hcjpeg._check(L.hcj_decode_batch(ctx._h, jp, lens, batch_n, mode, hcjpeg.FLAG_DEFAULT, op, caps, status))
            ctx.synchronize()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0) / e2e_steps
            assert all(status[i] == 0 for i in range(batch_n))
            got = np.ctypeslib.as_array(C.cast(op[0], C.POINTER(C.c_uint8)), shape=(caps[0],))
            assert np.array_equal(got, outs[0]), "e2e output differs from the resident-path output"
            result["e2e"] = {"value": world * mp_per_step / dt, "unit": "MP/s", "h2d_bytes_per_step": comp_bytes,
                             "d2h_bytes_per_step": out_bytes + 4 * batch_n, "ms_per_step": dt * 1e3, "steps": e2e_steps}
            L.hcj_host_free(pin_in)
            L.hcj_host_free(pin_out)
        cpu_items, cpu_kind = jpgs, "decode"
        metric = "decoded megapixels/sec (%dx%d %d baseline)" % (w, h, chroma)
    else:
        # ---- encode: frames in pinned host memory -> files in pinned host memory through hcj_encode_batch
        import ctypes as C

        L = hcjpeg.lib()
        frame_bytes = len(frames[0])
        cap = 1 << 20
        pin_in = L.hcj_host_alloc(frame_bytes * batch_n)
        pin_out = L.hcj_host_alloc(cap * batch_n)
        assert pin_in and pin_out
        fp = (C.c_void_p * batch_n)()
        op = (C.c_void_p * batch_n)()
        caps = (C.c_size_t * batch_n)(*([cap] * batch_n))
        lens = (C.c_size_t * batch_n)()
        status = (C.c_int * batch_n)()
        for i in range(batch_n):
            C.memmove(pin_in + i * frame_bytes, frames[i % unique], frame_bytes)
            fp[i], op[i] = pin_in + i * frame_bytes, pin_out + i * cap

        def run():
            hcjpeg._check(L.hcj_encode_batch(ctx._h, fp, batch_n, w, h, chroma, quality, ri, op, caps, lens, status))
            ms = C.c_float()
            hcjpeg._check(L.hcj_encode_last_device_ms(ctx._h, C.byref(ms)))
            return ms.value

        for _ in range(max(args.warmup, 3)):
            run()
        barrier()
        sampler.mark()
        dev_ms = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dev_ms.append(run())
        ctx.synchronize()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / args.steps
        clocks = sampler.stop()
        assert all(status[i] == 0 for i in range(batch_n))
        ref, st1 = ctx.encode_batch(frames[:1], w, h, chroma, quality, ri)
        got = bytes(np.ctypeslib.as_array(C.cast(op[0], C.POINTER(C.c_uint8)), shape=(lens[0],)))
        assert got == ref[0], "pinned-buffer output differs from the Python front-end's"
        kms = max_over_ranks(float(np.mean(dev_ms)))
        out_total = sum(lens[i] for i in range(batch_n))
        nblocks = hcjpeg.frame_info(ref[0]).nblocks * batch_n
        # FDCT+quantise stage: reads the frames, writes int16 coefficient blocks (in + 128 Nb)
        result.update(ms_per_step=kms, value=world * mp_per_step / (kms * 1e-3), gpu_launches=8 * args.steps, clocks=clocks,
                      roofline={"kernel": "encode pipeline (fdct_quant + bit lengths + pack + stuff)", "bound": "hbm",
                                "achieved": (frame_bytes * batch_n + 3 * 128 * nblocks + 3 * out_total) / (kms * 1e-3) / 1e9,
                                "peak": peaks()[0], "unit": "GB/s",
                                "frac": (frame_bytes * batch_n + 3 * 128 * nblocks + 3 * out_total) / (kms * 1e-3) / 1e9 / peaks()[0],
                                "traffic": None, "peak_source": peaks()[1],
                                "note": "algorithmic bytes: frames in + coefficient blocks written once and read twice + packed bits written, read, stuffed bytes written"},
                      e2e={"value": world * mp_per_step / dt, "unit": "MP/s", "h2d_bytes_per_step": frame_bytes * batch_n,
                           "d2h_bytes_per_step": out_total, "ms_per_step": dt * 1e3, "steps": args.steps})
        L.hcj_host_free(pin_in)
        L.hcj_host_free(pin_out)
        cpu_items, cpu_kind = frames, "encode"
        metric = "encoded megapixels/sec (%dx%d %d baseline)" % (w, h, chroma)

    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_arm(cpu_kind, cpu_items, w, h, chroma, quality, ri)
    line = {
        "metric": metric, "value": result["value"], "unit": "MP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": result["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int16/int32 (integer, bit-exact)", "data": "synthetic", "config": config,
        "clocks": result["clocks"], "e2e": result.get("e2e"), "gpu_launches": result["gpu_launches"],
        "roofline": result["roofline"], "cpu_baseline": cpu,
    }
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
