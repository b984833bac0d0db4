#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native OpenWebRX+ DSP hot path.

Metric (BASELINE.json): channel-MS/s = wideband input MS/s x concurrent client channels, measured on
BASELINE config 2: "Selector DDC + NFM/AM/USB demod: 10 MS/s wideband, 64 concurrent 12 kHz client
channels on 1xB200".  A step = one pass of the hot path (Shift -> FirDecimate -> FractionalDecimator ->
Bandpass -> Squelch -> demod -> AGC for all 64 channels) over one synthetic IQ block.

  python bench.py --gpus N --steps K --warmup W            # our arm (N > 1: under torchrun, weak scaling:
                                                           #   64 channels per GPU, IQ block broadcast over NCCL)
  python bench.py --impl reference ...                     # the CPU chain (oracle port) on the host cores

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = same metric through
the public host API (owrx_bank_feed from pinned host memory + audio read-back); `roofline` describes the
dominant kernel (K3: NCO mix + polyphase FIR decimation); `waterfall` reports the FftChain half of the
hot path (BASELINE config 1 shape) from its own timed loop; `cpu_baseline` the oracle on host cores.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 10_000_000           # config 2 wideband rate
OUT_RATE = 12000
CH_PER_GPU = 64
BLOCK = 1 << 24           # samples per step: 134 MB of complex64 > the 126 MB L2
WF_FS, WF_N, WF_FPS, WF_OV = 2_400_000, 4096, 9, 0.3     # config 1 (waterfall)
METRIC = "channel-MS/s (input MS/s x clients) + waterfall FFT frames/s"
WORKLOAD = "C2: Selector DDC + NFM/AM/USB demod, 10 MS/s wideband, %d x 12 kHz channels per GPU" % CH_PER_GPU


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
        except Exception:
            pass
    return 6650.0, 1965.0, "fallback"


def load_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r1_traffic.json);
    valid for this bench's block size / channel count only."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[kernel]
        if d.get("block_samples") == BLOCK and d.get("channels") == CH_PER_GPU:
            return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        pass
    return None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every few ms from a
    thread (the main thread sits in cudaStreamSynchronize with the GIL released); nvidia-smi -lms as fallback."""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index = index
        self.samples = []          # (sm_mhz, max_mhz, reason_mask)
        self.stop_flag = False
        self.thread = None
        self.proc = None
        self.lines = []

    def _nvml_loop(self):
        import pynvml
        h, mx = self.handle, self.max_mhz
        while not self.stop_flag:
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(sm), float(mx), int(mask), time.perf_counter()))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()                 # the slow part, done before any timed work
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """start of the timed region: the sampler has been running since before the warm-up (NVML start-up is slow), samples
        from here on are the ones reported"""
        self.t_mark = time.perf_counter()

    def stop(self):
        reasons = set()
        window = "timed region"
        if self.thread is not None:
            t_end = time.perf_counter()
            self.stop_flag = True
            self.thread.join(1.0)
            t0 = getattr(self, "t_mark", 0.0)
            inside = [x for x in self.samples if t0 <= x[3] <= t_end]
            if len(inside) < 3:
                # a timed region of a few ms holds too few polls: use every sample since the start of the warm-up (same load)
                inside, window = list(self.samples), "warm-up + timed region"
            for _, _, mask, _ in inside:
                for name, bit in self.REASONS:
                    if mask & bit:
                        reasons.add(name)
            sm = [x[0] for x in inside]
            mx = [x[1] for x in inside]
        elif self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            sm, mx = [], []
            for l in self.lines:
                f = [x.strip() for x in l.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source"], "samples": 0}
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm), "window": window}


def channel_plan(rank, n_ch):
    from openwebrx_b200.synth import carrier_plan
    return carrier_plan(64, FS, seed=20260101 + rank)[:n_ch] if n_ch <= 64 else carrier_plan(n_ch, FS, seed=20260101 + rank)


def synth_iq_torch(n, fs, carriers, device, seed=20260101):
    """Same signal model as openwebrx_b200.synth.make_iq, generated on the GPU (plumbing, untimed)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.empty(n, 2, device=device, dtype=torch.float32)
    chunk = 1 << 21
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        t = (torch.arange(s, e, device=device, dtype=torch.float64)) / fs
        re = 1e-3 * torch.randn(e - s, device=device, generator=g, dtype=torch.float32)
        im = 1e-3 * torch.randn(e - s, device=device, generator=g, dtype=torch.float32)
        for c in carriers:
            f, a, kind = c["offset"], c["amp"], c["kind"]
            if kind == "am":
                env = a * (1.0 + 0.5 * torch.cos(2 * np.pi * 1000.0 * t)) / 1.5
                ph = 2 * np.pi * ((f * t) % 1.0)
                re += (env * torch.cos(ph)).float(); im += (env * torch.sin(ph)).float()
            elif kind in ("nfm", "wfm"):
                ph = 2 * np.pi * ((f * t) % 1.0) + (2.5 if kind == "nfm" else 75.0) * torch.sin(2 * np.pi * 1000.0 * t)
                re += (a * torch.cos(ph)).float(); im += (a * torch.sin(ph)).float()
            else:
                for df in (700.0, 1900.0):
                    ph = 2 * np.pi * (((f + df) * t) % 1.0)
                    re += (0.5 * a * torch.cos(ph)).float(); im += (0.5 * a * torch.sin(ph)).float()
        x[s:e, 0] = re; x[s:e, 1] = im
    return x


def cpu_chain_rate(carriers, seconds_budget, n_samples, threads):
    """Times the oracle port of the per-client chain (one chain per client, one pass per stage — the
    reference's structure) on `threads` host threads.  Returns (channel-MS/s, channels run, wall s)."""
    import oracle
    from openwebrx_b200.synth import BANDPASS, make_iq
    kind = {"nfm": oracle.DEMOD_NFM, "am": oracle.DEMOD_AM, "usb": oracle.DEMOD_SSB}
    iq = make_iq(n_samples, FS, carriers[:8], seed=1)
    oracle.lib()

    def one(c):
        oracle.client_chain_run(iq, FS, OUT_RATE, c["offset"], BANDPASS[c["kind"]], kind[c["kind"]], fast_shift=True)

    t0 = time.perf_counter(); one(carriers[0]); t1 = time.perf_counter() - t0
    per_thread = max(1, int(seconds_budget / max(t1, 1e-3)))
    done = [0] * threads

    def worker(i):
        for k in range(per_thread):
            one(carriers[(i + k) % len(carriers)])
            done[i] += 1

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    wall = time.perf_counter() - t0
    total = sum(done)
    return total * n_samples / wall / 1e6, total, wall


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  pycsdr/libcsdr are not in the
    reference tree and cannot be built, so this times the oracle port (cpu_baseline.kind
    "port") with all host threads, on the same config / metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    carriers = channel_plan(0, CH_PER_GPU)
    n_samples = 1 << 20
    vals, ms = [], []
    for step in range(args.warmup + args.steps):
        v, total, wall = cpu_chain_rate(carriers, 1.0, n_samples, threads)
        if step >= args.warmup:
            vals.append(v); ms.append(wall * 1e3)
    value = float(np.mean(vals)) if vals else 0.0
    sample = "%d threads x oracle client chains (10 MS/s -> 12 kHz NFM/AM/USB) over %d-sample records, ~1 s per step" % (threads, n_samples)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "channel-MS/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(ms)) if ms else None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "cpu": True},
            "cpu_baseline": {"value": value, "unit": "channel-MS/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "channel-MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from openwebrx_b200 import ChannelBank, Waterfall, _native as N, fftchain_params
    from openwebrx_b200.synth import BANDPASS
    from openwebrx_b200.sharding import MulticastHop, broadcast_block

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the hop runs beside the DSP kernels: NCCL's copy CTAs take SMs from the latency-bound low-rate stages.  16 channels
        # carry the 134 MB block in ~0.26 ms (hidden behind the 0.33 ms DSP pass) with half the CTAs of the default; measured
        # on 2 B200 (tools/nccl_channels_sweep.sh): 4 -> 0.95, 8 -> 0.52, 16 -> 0.39, default -> 0.45 ms per step
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "16")
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, sm_max, peak_src = load_peaks()

    # ---- channels: 64 per GPU, channel c tunes to carrier c (1/3 AM, 1/3 NFM, 1/3 USB)
    carriers = channel_plan(0, CH_PER_GPU)          # the wideband signal (same on every rank: it is broadcast)
    my_plan = carriers                              # each rank tunes its own 64 clients to those carriers
    bank = ChannelBank(FS, device=local)
    chans = []
    for i, c in enumerate(my_plan):
        off = c["offset"] + (rank * 7) % 50          # ranks tune slightly differently: no shared work
        chans.append(bank.add_channel(OUT_RATE, demod=c["kind"], offset=off, bandpass=BANDPASS[c["kind"]]))
    D, T = 833, 22223
    n_k = (BLOCK - T) // D + 1
    consumed = n_k * D

    # ---- synthetic wideband block, resident in HBM (rank 0 generates; others receive it by broadcast)
    iq = synth_iq_torch(BLOCK, FS, carriers, dev) if rank == 0 else torch.empty(BLOCK, 2, device=dev, dtype=torch.float32)
    if world > 1:
        dist.broadcast(iq, 0)
    iq_src = iq
    # the hop: NVSwitch multicast (one multimem.st pass on the ingest GPU, see openwebrx_b200/sharding.py) when the GPUs
    # support it, else an NCCL broadcast; OWRX_HOP=nccl forces the latter
    hop, hop_kind = None, "none"
    if world > 1:
        hop_kind = "nccl-broadcast" if os.environ.get("OWRX_HOP") != "none" else "NONE (diagnostic run: not a valid multi-GPU number)"
        # measured on 2 and 8 B200 (tools/hop_sweep.sh): both hops hide behind the DSP pass; NCCL's is the faster end to end
        # (0.49 vs 0.61 ms per step at 8 GPUs: the multicast protocol's two cross-GPU barriers per block cost more than
        # NCCL's copy kernels), so it is the default and OWRX_HOP=multicast selects the NVLS form
        if os.environ.get("OWRX_HOP", "nccl") == "multicast":
            try:
                hop = MulticastHop(2 * BLOCK, dev)
                hop_kind = "nvls-multicast"
            except Exception as e:                       # no NVLS on this box / torch build: keep NCCL
                print("[bench] multicast hop unavailable (%s: %s); using NCCL broadcast" % (type(e).__name__, e), file=sys.stderr)
                hop = None
        flag = torch.tensor([1 if hop is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # all ranks or none
        if int(flag.item()) == 0:
            hop, hop_kind = None, "nccl-broadcast"
    bcast_buf = [torch.empty_like(iq), torch.empty_like(iq)] if (world > 1 and hop is None) else [iq]
    iq_alt = [iq, iq.clone()] if world == 1 else None
    # a created (non-default) stream: the C ABI treats a NULL handle as "use the object's own stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    pending = [None]

    def issue_broadcast(i):
        # the hop: rank 0's block reaches every GPU over NVLink (NCCL broadcast).  Issued one block ahead on NCCL's own
        # stream, so the transfer of block i+1 overlaps the K3 pass of block i (double-buffered).
        # rank 0 sends its resident block as it is (inputs are resident in HBM when the timed region starts); the others
        # receive into alternating buffers
        buf = iq_src if rank == 0 else bcast_buf[i & 1]
        pending[0] = broadcast_block(buf, 0, async_op=True)

    sent = [0]

    def step(i):
        if hop is not None:
            # block i+1 crosses the switch on the hop stream while block i is processed; buffers alternate
            while sent[0] <= i + 1:
                hop.send(sent[0], iq_src if rank == 0 else None)
                sent[0] += 1
            buf = hop.recv(i, stream)
            bank.process_device(buf, BLOCK, stream=sp)
            hop.release(i, stream)
        elif world > 1 and os.environ.get("OWRX_HOP") == "none":
            bank.process_device(iq_src, BLOCK, stream=sp)    # diagnostic only: no hop, every rank reads its resident copy
        elif world > 1:
            if pending[0] is None:
                issue_broadcast(i)
            pending[0].wait()                            # current stream waits for block i
            issue_broadcast(i + 1)
            bank.process_device(iq_src if rank == 0 else bcast_buf[i & 1], BLOCK, stream=sp)
        else:
            # two resident blocks, alternated: 268 MB of input between two reads of the same bytes (L2 is 126 MB)
            bank.process_device(iq_alt[i & 1], BLOCK, stream=sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    bank.set_pipelined(True)       # low-rate stages of block i overlap the K3 pass of block i+1 (side stream)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()             # NVML start-up takes longer than a short timed region: poll from the warm-up on
    for i in range(args.warmup):
        step(i)
    bank.join(sp)
    barrier()
    if hop is not None:
        # the hop delivers the ingest rank's block bit for bit: compare a checksum of the last received buffer
        chk = hop.bufs[(args.warmup - 1) & 1].double().sum().reshape(1)
        ref = iq_src.double().sum().reshape(1) if rank == 0 else torch.zeros(1, device=dev, dtype=torch.float64)
        dist.broadcast(ref, 0)
        assert torch.equal(chk, ref), "multicast hop delivered a different block"
    bank.profile(True)
    bank.profile_read(reset=True)
    launches0 = N.lib.owrx_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.mark()
    ev0.record(stream)
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps     # CPU time to enqueue one step (diagnostic)
    bank.join(sp)                  # the timed region ends when the last block's audio is complete
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = N.lib.owrx_launch_count() - launches0
    prof = bank.profile_read_ex(reset=True)
    bank.profile(False)
    fastconv = prof["fc_contract"][1] > 0
    fir_form = bank.fir_form()
    tc = fir_form == "fastconv_tc"
    dom = "fc_contract" if fastconv else "k3_direct"
    k3_ms, k3_n = prof[dom]
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * CH_PER_GPU * consumed / (ms_step * 1e-3) / 1e6

    # ---- e2e: public host API, pinned host input, H2D + audio D2H inside the timed region
    h_iq = torch.empty(BLOCK, 2, dtype=torch.float32).pin_memory()
    h_iq.copy_(iq)
    bank2 = ChannelBank(FS, device=local)
    ch2 = [bank2.add_channel(OUT_RATE, demod=c["kind"], offset=c["offset"], bandpass=BANDPASS[c["kind"]]) for c in my_plan]
    hp = h_iq.data_ptr()

    audio_buf = np.empty((len(ch2), 1 << 15), np.float32)      # ~20 160 samples per channel and block

    up_count = [0]
    if world > 1 and hop is None:
        up_stream = torch.cuda.Stream(device=dev)
        up_buf = [torch.empty_like(iq), torch.empty_like(iq)]
        up_done = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_step():
        """one block through the public host API: H2D of the block from pinned memory, the whole chain, audio D2H and
        the read of every channel's audio into a caller buffer (owrx_bank_read_audio_all)"""
        if world > 1:
            # rank 0 uploads, NCCL carries the block to the other GPUs, every rank returns its audio to the host
            if hop is not None:
                if rank == 0:
                    iq.copy_(h_iq, non_blocking=True)
                j = sent[0]
                sent[0] += 1
                hop.stream.wait_stream(stream)           # the upload is on the bench stream
                hop.send(j, iq if rank == 0 else None)
                buf = hop.recv(j, stream)
                bank2.process_device(buf, BLOCK, stream=sp)
                hop.release(j, stream)
            else:
                # streaming: rank 0 uploads block j+1 (copy stream) while block j crosses NVLink and is processed
                j = up_count[0]
                up_count[0] += 1
                if rank == 0:
                    if j == 0:
                        with torch.cuda.stream(up_stream):
                            up_buf[0].copy_(h_iq, non_blocking=True)
                            up_done[0].record(up_stream)
                    stream.wait_event(up_done[j & 1])
                    up_stream.wait_stream(stream)                # buffer (j+1)&1 was read by the hop of block j-1
                    with torch.cuda.stream(up_stream):
                        up_buf[(j + 1) & 1].copy_(h_iq, non_blocking=True)
                        up_done[(j + 1) & 1].record(up_stream)
                buf = up_buf[j & 1]
                broadcast_block(buf, 0)
                bank2.process_device(buf, BLOCK, stream=sp)
            bank2.drain()
            return sum(bank2.read_audio_all(ch2, audio_buf))
        bank2.feed_ptr(hp, BLOCK)
        return sum(bank2.read_audio_all(ch2, audio_buf))

    def e2e_loop(steps):
        got = 0
        for i in range(steps):
            got += e2e_step()
        if world == 1:
            bank2.flush()                                 # streaming mode: the last block's final outputs
            got += sum(bank2.read_audio_all(ch2, audio_buf))
        return got

    # the host API in its streaming mode (owrx_bank_set_deferred_drain): a feed only enqueues and returns, the kernel tail, D2H
    # and queue hand-over of block i run under the upload of block i+1 (uploads queue back to back: PCIe never idles); every
    # block is still uploaded from pinned host memory and every block's audio is read back inside the timed region (the last
    # one after a flush)
    if world == 1:
        bank2.set_deferred_drain(True)
    e2e_loop(max(1, min(args.warmup, 3)))
    barrier()
    e2e_steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    n_audio = e2e_loop(e2e_steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_sync = None
    if world == 1:
        # the same with the synchronous default (each feed returns with its outputs in the host queues)
        bank2.set_deferred_drain(False)
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        ms_sync = (time.perf_counter() - t0) * 1e3 / e2e_steps
        e2e_sync = {"value": CH_PER_GPU * (BLOCK // D) * D / (ms_sync * 1e-3) / 1e6, "unit": "channel-MS/s", "ms_per_step": ms_sync,
                    "note": "owrx_bank_feed in its default synchronous mode"}
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_consumed = (BLOCK // D) * D                   # the streaming path carries the FIR tail between blocks
    e2e_value = world * CH_PER_GPU * e2e_consumed / (e2e_ms * 1e-3) / 1e6
    d2h = (n_audio // e2e_steps) * 4

    # ---- the same end-to-end step for a source that delivers complex int16 (the reference converts such
    # sources on the CPU, owrx/source/fifi_sdr.py:27-28): half the PCIe bytes, Convert on the GPU.  Reported beside e2e.
    e2e_cs16 = None
    if world == 1:
        h16 = torch.empty(BLOCK, 2, dtype=torch.int16).pin_memory()
        h16.copy_((iq.clamp(-1, 1) * 32767.0).to(torch.int16))
        bank3 = ChannelBank(FS, device=local)
        ch3 = [bank3.add_channel(OUT_RATE, demod=c["kind"], offset=c["offset"], bandpass=BANDPASS[c["kind"]]) for c in my_plan]
        hp16 = h16.data_ptr()
        bank3.set_deferred_drain(True)
        for i in range(3):
            bank3.feed_ptr(hp16, BLOCK, fmt="cs16")
            bank3.read_audio_all(ch3, audio_buf)
        bank3.flush(); bank3.read_audio_all(ch3, audio_buf)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            bank3.feed_ptr(hp16, BLOCK, fmt="cs16")
            bank3.read_audio_all(ch3, audio_buf)
        bank3.flush(); bank3.read_audio_all(ch3, audio_buf)
        torch.cuda.synchronize()
        ms16 = (time.perf_counter() - t0) * 1e3 / e2e_steps
        e2e_cs16 = {"value": CH_PER_GPU * e2e_consumed / (ms16 * 1e-3) / 1e6, "unit": "channel-MS/s", "h2d_bytes_per_step": BLOCK * 4,
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms16,
                    "note": "same step fed with complex int16 host samples (owrx_bank_feed_fmt): Convert runs on the GPU"}
        bank3.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- waterfall half of the hot path (config 1 shape), own timed loop on this GPU
    wf_stats = bench_waterfall(torch, dev, hbm_peak)

    # ---- roofline of the dominant kernel: the fast-convolution contraction (K3F) or, in direct mode, K3
    k3_avg_s = (k3_ms / max(k3_n, 1)) * 1e-3
    f_obs = (clk or {}).get("sm_mhz") or sm_max
    fp32_peak = 148 * 128 * 2 * f_obs * 1e6 / 1e12
    if fastconv:
        M, P = 256, -(-T // D)
        Kb, Dp = M - P + 1, -(-D // 32) * 32
        B = -(-n_k // Kb)                                         # overlap-save blocks per launch
        flops = 8.0 * M * B * Dp * CH_PER_GPU                     # complex MAC = 4 FMA per (bin, block, branch, channel)
        if tc:
            # operands of the tensor-core contraction, each moved once: branch spectra and table as 3 bf16 terms per float
            # (12 B per complex entry), Z as complex float32 (8 B)
            algo_bytes = 12.0 * M * B * Dp + 12.0 * M * Dp * CH_PER_GPU + 8.0 * M * B * CH_PER_GPU
            kname = ("fc_contract_tc_kernel (K3F: per-channel spectral contraction over the %d polyphase branches, %d blocks x %d ch, "
                     "tcgen05 bf16x3 -> FP32 in TMEM)" % (D, B, CH_PER_GPU))
            knote = ("HBM bound (ncu: DRAM 63 % of its peak, tensor pipe active 14 % of cycles; issued MMA FLOPs in roofline_tensor); timed "
                     "inside the three-stream pipeline (standalone ncu capture: profiles/r1_fc_contract_tc.md, 83.5 us); the shared forward "
                     "FFTs (fc_forward) and the inverse FFT + rotation (fc_inverse) are in stages_ms")
            tkey = "fc_contract_tc_kernel"
        else:
            # operands of the contraction, each moved once: F (16 B, packed-FMA layout), table (8 B), Z (8 B)
            algo_bytes = 16.0 * M * B * Dp + 8.0 * M * Dp * CH_PER_GPU + 8.0 * M * B * CH_PER_GPU
            kname = "fc_contract_kernel (K3F: per-channel spectral contraction over the %d polyphase branches, %d blocks x %d ch)" % (D, B, CH_PER_GPU)
            knote = ("FP32-FMA bound (%.0f FLOP per operand byte): see roofline_fp32; the shared forward FFTs (fc_forward) and the inverse "
                     "FFT + rotation (fc_inverse) are in stages_ms" % (flops / algo_bytes))
            tkey = "fc_contract_kernel"
    else:
        algo_bytes = 8.0 * BLOCK + 8.0 * CH_PER_GPU * n_k        # IQ read once + complex IF written (per launch)
        flops = CH_PER_GPU * float(consumed) * (8.0 + 4.0 * T / D)    # C*Nin*(8 + 4T/D)
        kname = "fir_decimate_kernel (K3: NCO mix + polyphase FIR decimate, 64 ch)"
        knote = "direct-form DDC is FP32-FMA bound by construction (SURVEY 8d): see roofline_fp32"
        tkey = "fir_decimate_kernel"
    achieved = algo_bytes / k3_avg_s / 1e9 if k3_avg_s > 0 else 0.0
    fp32_ach = flops / k3_avg_s / 1e12 if k3_avg_s > 0 else 0.0
    stages = {k: v[0] / v[1] for k, v in prof.items() if v[1]}
    tensor = None
    if fastconv and tc:
        # issued bf16 MMA work of the same launch: 6 partial products x 2 accumulator halves per (128-row tile, 128 columns, 16 branches)
        mma_flops = 6.0 * 2.0 * (2.0 * 128 * 128 * 16) * (Dp // 16) * M * (-(-B // 128)) * (CH_PER_GPU // 64)
        try:
            bf16_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            bf16_peak = 1646.0
        t_ach = mma_flops / k3_avg_s / 1e12 if k3_avg_s > 0 else 0.0
        tensor = {"achieved": t_ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": t_ach / bf16_peak,
                  "note": "issued bf16 MMA FLOPs incl. the padded rows of the 128-row tile; the kernel is HBM bound (roofline), not tensor bound; "
                          "roofline_fp32.achieved is the same launch in FP32-equivalent FLOPs"}
    # the same work expressed in the reference's own terms (direct form: SURVEY 8d) for comparison across forms
    direct_equiv = CH_PER_GPU * float(consumed) * (8.0 + 4.0 * T / D) / (ms_step * 1e-3) / 1e12

    # ---- CPU baseline beside it (bounded sample, rank 0 only, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, total, wall = cpu_chain_rate(carriers, 2.0, 1 << 20, threads)
        cpu = {"value": v, "unit": "channel-MS/s", "cores": threads, "kind": "port",
               "sample": "%d oracle client chains (C2 shape) over 2^20-sample records on %d threads, %.1f s wall" % (total, threads, wall)}

    line = {
        "metric": METRIC, "value": value, "unit": "channel-MS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "channels_total": world * CH_PER_GPU, "block_samples": BLOCK, "decimation": D, "fir_taps": T,
                   "l2": "two resident 134 MB input blocks alternate (268 MB between re-reads > 126 MB L2); no flush needed", "parallelism": "channels sharded x%d, IQ block hop: %s" % (world, hop_kind) if world > 1 else "1 GPU",
                   "realtime_factor": value / (FS / 1e6 * CH_PER_GPU * world)},
        "e2e": {"value": e2e_value, "unit": "channel-MS/s", "h2d_bytes_per_step": BLOCK * 8, "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms,
                "mode": "streaming (owrx_bank_set_deferred_drain): block i's outputs are drained under block i+1's upload; all "
                        "blocks' audio read inside the timed region (last block after owrx_bank_flush)" if world == 1 else "streaming: H2D of block j+1 on rank 0 under the hop + owrx_bank_process_device + drain + reads of block j"},
        "e2e_sync": e2e_sync,
        "e2e_cs16": e2e_cs16,
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "clocks": clk,
        "roofline": {"kernel": kname, "bound": "hbm", "achieved": achieved,
                     "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": load_traffic(tkey), "peak_source": peak_src,
                     "kernel_ms": k3_avg_s * 1e3, "kernel_share_of_step": (k3_ms / args.steps) / ms_step if ms_step > 0 else None,
                     "algorithmic_bytes": algo_bytes, "note": knote},
        "roofline_fp32": {"achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_ach / fp32_peak if fp32_peak else None,
                          "peak_def": "148 SM x 128 lanes x 2 x observed SM clock (%.0f MHz)" % f_obs,
                          "direct_form_equivalent_tflops": direct_equiv},
        "roofline_tensor": tensor,
        "fir_form": fir_form,
        "stages_ms": stages,
        "waterfall": wf_stats,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_waterfall(torch, dev, hbm_peak, lines=592, steps=5):
    from openwebrx_b200 import Waterfall, fftchain_params
    avg, every_n = fftchain_params(WF_FS, WF_N, WF_OV, WF_FPS)
    n = every_n * avg * lines + WF_N                      # 592 lines = 1.27 GB of IQ, far beyond L2
    g = torch.Generator(device=dev); g.manual_seed(7)
    iq = 1e-3 * torch.randn(n, 2, device=dev, generator=g, dtype=torch.float32)
    tt = torch.arange(n, device=dev, dtype=torch.float32)
    iq[:, 0] += 0.3 * torch.cos(0.7 * tt); iq[:, 1] += 0.3 * torch.sin(0.7 * tt)
    del tt
    wf = Waterfall(WF_FS, WF_N, WF_OV, WF_FPS, "adpcm", device=dev.index or 0)
    out = torch.empty(lines * wf.line_bytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    assert st.cuda_stream != 0
    torch.cuda.synchronize()
    wf.set_pipelined(True)        # ADPCM of batch i on the side stream beside the FFT pass of batch i+1
    for _ in range(3):
        wf.process_device(iq, n, out, out.numel(), stream=st.cuda_stream)
    wf.join(st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        got = wf.process_device(iq, n, out, out.numel(), stream=st.cuda_stream)
    wf.join(st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    unique = 8.0 * ((avg - 1) * every_n + WF_N) + wf.line_bytes      # 8*U + out per line
    gbs = unique * got / (ms * 1e-3) / 1e9
    return {"workload": "C1: 2.4 MS/s, 4096-pt, 9 fps, overlap 0.3 -> avg 93, hop 2867, ADPCM", "lines_per_s": got / (ms * 1e-3),
            "ffts_per_s": got * avg / (ms * 1e-3), "realtime_factor": got / (ms * 1e-3) / 9.0, "ms_per_batch": ms, "lines_per_batch": int(got),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                         "note": "whole FftChain (fft + finalize + adpcm launches) vs algorithmic bytes"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE JSON line: anything a library prints on fd 1 meanwhile (NCCL's version banner under
    # NCCL_DEBUG=VERSION, ...) is sent to stderr, and the line is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    sys.stdout.write(out.getvalue())
    sys.stdout.flush()


if __name__ == "__main__":
    main()
