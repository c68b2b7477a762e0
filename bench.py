#!/usr/bin/env python3
"""Benchmark of the VPD student hot path (BASELINE.json metric: student train frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one distillation training step of the ResNet-34 student on a batch of
256 synthetic 128x128 RGB+flow crops per GPU (BASELINE.json configs[1]): K1 batch
assembly from device-resident uint8 pools -> train-mode forward -> sum-MSE loss ->
backward -> [NCCL all-reduce SUM of the gradient arena when N > 1] -> fused AdamW.

  value     device-timed frames/s over all ranks (CUDA events, max over ranks), inputs already
            in HBM; W warm-up steps, then >= --preheat-s seconds of the same steps (so the
            clocks sampled during the timed region are steady-state), then EXACTLY K timed steps;
  e2e       the same metric through the reference-facing call `ModelTrainer.epoch` fed with
            pinned HOST batches in the reference loader's own format, fp32 {'img','emb'}
            (H2D copies and the loss read-back inside the timed region); `e2e.u8_batches` is
            the same with raw uint8 crop batches (4x fewer bytes, K1 on the device);
  roofline  the implicit-GEMM conv family (forward + data-gradient launches): algorithmic FLOPs
            / kernel time. Kernel times are hardware start/end timestamps of every launch (CUPTI
            activity records through torch.profiler) of a few extra steps run SERIALISED after
            the timed region (eager, one stream - as replayed in the timed region every kernel
            starts early under programmatic dependent launch and waits for its predecessor, so
            as-run durations overlap); the CUDA-event intervals around the same launches are
            kept beside them (`achieved_events` - they include the gaps between launches). `frac` is against the measured bf16 peak that matches
            the conditions of the timed region (sustained when the GPU sits at its power cap or
            below 90 % of its maximum clock, burst otherwise); both fractions are printed. `kernels` has the other families with their own
            rooflines (K1 assembly and AdamW against measured HBM bandwidth);
  dp_check  (N > 1) one extra checked step after the timed region: per all-reduce bucket, the
            trainer's gradient arena vs the sum over ranks of every rank's own gradients;
  configs   the other BASELINE.json configs: forward at batch 32 (config 1, with the reference's
            CPU `embed` beside it), keypoint-teacher training at n = 4096 (config 4), corpus
            feature extraction to per-video pickles (config 5, scaled - see `what`);
  cpu_baseline  the UNMODIFIED reference's `ModelTrainer.epoch` (oracle/_ref, see
            oracle/build_ref.py) on the host cores: 2 steps of the same batch-256 workload.
`--impl reference` times only that CPU path, at batch 256, for the full --steps.
"""
import argparse
import ctypes
import json
import math
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 256
IMG = 128
EMB = 32
POOL = 4096
TRAIN_FLOP_PER_FRAME = 7.2029e9     # fwd 2.4438 + bwd 4.7591 GFLOP
FWD_FLOP_PER_FRAME = 2.4438e9
K1_BYTES_PER_FRAME = 81920 + 291584   # uint8 rgb + flow-xy read, bf16 stem layout written
ADAMW_BYTES_PER_PARAM = 28
METRIC = 'vpd_student_train_frames_per_s'
UNIT = 'frames/s'


def get_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', type=str, default='native', choices=['native', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--preheat-s', type=float, default=2.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-configs', action='store_true')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fp:
            p = json.load(fp)
        return {'hbm_gbs': p['hbm_gbs'], 'tflops': p['bf16_tflops_sustained'],
                'tflops_burst': p['bf16_tflops'], 'source': 'MEASURED_PEAKS.json'}
    return {'hbm_gbs': 6650.0, 'tflops': 1400.0, 'tflops_burst': 1590.0,
            'source': 'fallback (B200_PROFILING.md)'}


# --------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref or /root/reference), else the oracle port
# --------------------------------------------------------------------------------------
def _train_batch(batch):
    from oracle import assemble_ref
    from vpd_b200 import synth
    rgb, flow = synth.crops(batch, seed=1)
    teach = synth.teacher(batch, seed=3, emb_dim=EMB, motion=True)
    fl = synth.flips(batch, seed=2)
    return assemble_ref.train_batch(rgb.numpy(), flow.numpy(), teach.numpy(), fl.numpy(),
                                    *synth.FS_MEAN_STD)


def cpu_trainer():
    """-> (kind, step_fn(batches) running K train steps through `epoch`, description)"""
    from oracle import ref_shim
    torch.manual_seed(0)
    if ref_shim.available():
        ns = ref_shim.load()
        enc = ns.RGBF_EmbeddingModel('resnet34', EMB, True, 'cpu')
        tr = ns.ModelTrainer(enc, True)
        opt, scaler = tr.get_optimizer(5e-4)
        what = ("the unmodified reference's train_vpd_model.ModelTrainer.epoch (fp32, torch CPU; "
                "sources: {})".format('oracle/_ref copy' if ref_shim.source() == 'copy'
                                      else ref_shim.REFERENCE_DIR))
        return 'reference', (lambda bs: tr.epoch(bs, optimizer=opt, scaler=scaler)), what
    from oracle import student_ref
    otr = student_ref.OracleTrainer(student_ref.init_encoder_state('resnet34', EMB, True),
                                    student_ref.init_decoder_state(EMB), lr=5e-4)

    def run(bs):
        tot = 0.0
        for b in bs:
            tot += otr.step(b['img'], b['emb'])
        return tot / sum(b['img'].shape[0] for b in bs)
    return 'port', run, 'oracle port of ModelTrainer.epoch (fp32, torch CPU; no reference copy found)'


def cpu_train_fps(steps, warmup, batch=BATCH, max_seconds=None):
    """frames/s of the reference's training step on the host cores, batch `batch` (never
    shrunk); with `max_seconds` the number of TIMED steps is cut (and reported) instead."""
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    kind, run, what = cpu_trainer()
    img, tgt = _train_batch(batch)
    b = {'img': img, 'emb': tgt}
    t0 = time.perf_counter()
    run([b] * max(1, warmup))
    per_step = (time.perf_counter() - t0) / max(1, warmup)
    ran = steps
    if max_seconds is not None:
        ran = max(1, min(steps, int(max_seconds / per_step)))
    t0 = time.perf_counter()
    loss = run([b] * ran)
    total = time.perf_counter() - t0
    return {'value': batch * ran / total, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': '{} train steps of {} frames after {} warm-up, {} threads; {}'.format(
                ran, batch, max(1, warmup), cores, what),
            'ms_per_step': 1e3 * total / ran, 'steps_run': ran, 'batch': batch,
            'loss_per_frame': loss}


def cpu_embed_fps(batch=32, reps=3):
    """config 1 on the host: the reference's RGBF_EmbeddingModel.embed, batch 32"""
    from oracle import ref_shim, assemble_ref, student_ref
    from vpd_b200 import synth
    import numpy as np
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    rgb, flow = synth.crops(batch, seed=0)
    x = assemble_ref.train_batch(rgb.numpy(), flow.numpy(), np.zeros((batch, 2, 2 * EMB), np.float32),
                                 np.zeros(batch, np.uint8), *synth.FS_MEAN_STD)[0].numpy()
    torch.manual_seed(0)
    if ref_shim.available():
        ns = ref_shim.load()
        enc = ns.RGBF_EmbeddingModel('resnet34', EMB, True, 'cpu')
        kind, fn = 'reference', (lambda: enc.embed(x))
    else:
        sd = student_ref.init_encoder_state('resnet34', EMB, True)
        kind, fn = 'port', (lambda: student_ref.embed(sd, torch.from_numpy(x)))
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    return {'value': batch / dt, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'ms_per_call': dt * 1e3, 'sample': '{} embed() calls of {} frames'.format(reps, batch)}, x


def run_reference(args, rank):
    if rank != 0:
        return
    res = cpu_train_fps(args.steps, max(1, min(args.warmup, 2)), batch=args.batch, max_seconds=420.0)
    cfg = workload_config(args.gpus, args.batch)
    if res['steps_run'] != args.steps:
        cfg['note'] = ('timed {} of the {} requested steps at the full batch (a CPU step takes {:.1f} s): '
                       'the batch is never shrunk'.format(res['steps_run'], args.steps,
                                                          res['ms_per_step'] / 1e3))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': res['value'], 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': res['steps_run'], 'warmup': args.warmup,
        'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic', 'config': cfg,
        'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': res['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'loss_per_frame': res['loss_per_frame'],
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, batch):
    return {'workload': 'VPD student distillation training step (ResNet-34, 5x128x128 RGB+flow, '
                        'emb 32, --motion decoder, AdamW lr 5e-4), batch {} per GPU'.format(batch),
            'per_gpu_batch': batch, 'global_batch': batch * n_gpus, 'img_dim': IMG,
            'parallelism': 'dp{}'.format(n_gpus),
            'l2': 'inputs larger than L2: {} uint8 crops ({} MB) sampled at random each step'.format(
                POOL, POOL * IMG * IMG * 6 // (1 << 20))}


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '10'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def mark(self):
        """Samples taken from now on belong to the timed region."""
        self.first = len(self.rows)

    def _summary(self, rows):
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm),
                'reasons': sorted(reasons)}

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = self.rows[self.first:]
        if len(rows) < 2:
            rows = self.rows[-4:]
        out = self._summary(rows)
        pre = self._summary(self.rows[:self.first])
        out['preheat'] = {k: pre[k] for k in ('sm_mhz', 'power_w_max', 'samples', 'reasons')}
        return out


# --------------------------------------------------------------------------------------
# kernel timing
# --------------------------------------------------------------------------------------
KINDS = ['conv_fwd', 'bn_relu_pool_fwd', 'conv_dgrad', 'conv_wgrad', 'bn_relu_pool_bwd',
         'head_loss', 'pack_convert', 'other']


def conv_flops(batch):
    """Algorithmic FLOPs per step of the three conv kernel families (2*MACs), from the
    layer table: returns (fwd, dgrad, wgrad)."""
    from vpd_b200 import init as vinit
    fwd = 2.0 * batch * 64 * 64 * 64 * 5 * 49
    dgrad = 0.0
    h = IMG // 4
    for _, cin, cout, stride, ds in vinit.blocks('resnet34'):
        ho = h // stride
        c1 = 2.0 * batch * ho * ho * cout * cin * 9
        c2 = 2.0 * batch * ho * ho * cout * cout * 9
        d = 2.0 * batch * ho * ho * cout * cin if ds else 0.0
        fwd += c1 + c2 + d
        dgrad += c1 + c2 + d
        h = ho
    return fwd, dgrad, fwd


def classify_kernel(name):
    """CUPTI kernel name (demangled) of the TRAINING step -> family"""
    m = re.search(r'vpd::(\w+)(?:<([^>]*)>)?', name)
    if not m:
        return 'other', name
    base, targs = m.group(1), [a.strip() for a in (m.group(2) or '').split(',')]
    short = base + ('<' + ','.join(targs) + '>' if m.group(2) else '')
    if base in ('conv_wgrad_kernel', 'conv_wgrad_halo_kernel'):
        return 'conv_wgrad', short
    if base in ('conv_igemm_kernel', 'conv3x3_halo_kernel', 'conv3x3_halo_stream_kernel'):
        # MODE (1 = training forward, 0 / 2 / 3 = data gradients in the train step) is the last
        # template argument, except for conv3x3_halo_kernel<CHUNKS, MODE, NTAPS>
        mode = targs[1] if base == 'conv3x3_halo_kernel' else targs[-1]
        return ('conv_fwd' if mode == '1' else 'conv_dgrad'), short
    if base in ('bn_apply_kernel', 'bn_pool_kernel', 'channel_stats_kernel'):
        return 'bn_relu_pool_fwd', short
    if base in ('bn_bwd_kernel', 'stem_bwd_reduce_kernel', 'stem_bwd_reduce_sel_kernel',
                'stem_bwd_apply_kernel'):
        return 'bn_relu_pool_bwd', short
    if base in ('head_kernel', 'head_wgrad_kernel'):
        return 'head_loss', short
    if base.startswith('assemble'):
        return 'assemble', short
    if base in ('adamw_kernel', 'adamw_mirror_kernel', 'sgd_kernel'):
        return 'adamw', short
    if base in ('mirror_weights_kernel', 'cast_weights_kernel', 'nchw_to_pad8_kernel'):
        return 'pack_convert', short
    return 'other', short


def cupti_kernel_times(step_fn, nsteps):
    """Run `nsteps` steps under CUPTI activity tracing (torch.profiler, CUDA activities only)
    -> {family: {'us': total, 'n': launches, 'by_kernel': {name: [n, us]}}} per nsteps, or None
    when the profiler is not usable. Durations are the hardware start/end timestamps of each
    launch. The caller runs the steps SERIALISED (one stream, an event record between
    launches): as replayed in the timed region every kernel starts early under programmatic
    dependent launch and spins in griddepcontrol.wait until its predecessor drains, and the
    weight gradients share the SMs with the main chain, so as-run durations overlap."""
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(nsteps):
                step_fn(i)
            torch.cuda.synchronize()
        fam = {}
        for e in prof.events():
            if e.device_type != torch.autograd.DeviceType.CUDA:
                continue
            name = e.name
            if name.startswith(('Memcpy', 'Memset')) or 'vpd::' not in name:
                continue
            dur = e.time_range.elapsed_us()
            f, short = classify_kernel(name)
            d = fam.setdefault(f, {'us': 0.0, 'n': 0, 'by_kernel': {}})
            d['us'] += dur
            d['n'] += 1
            k = d['by_kernel'].setdefault(short, [0, 0.0])
            k[0] += 1
            k[1] += dur
        return fam if fam else None
    except Exception as exc:      # profiler unavailable on this box: fall back to events
        sys.stderr.write('bench: CUPTI kernel timing unavailable ({}: {})\n'.format(
            type(exc).__name__, exc))
        return None


def latest_traffic():
    """DRAM bytes per launch of the family's heaviest kernel from the newest committed
    `ncu --set full` capture (profiles/r*_traffic.json; not measurable without the profiler)."""
    pdir = os.path.join(ROOT, 'profiles')
    cands = sorted(f for f in os.listdir(pdir) if re.match(r'r\d+\w*_traffic\.json$', f))
    if not cands:
        return None, None
    with open(os.path.join(pdir, cands[-1])) as fp:
        tj = json.load(fp)
    of = {k: tj[k] for k in ('kernel', 'algorithmic_bytes_per_launch', 'tensor_pipe_active_pct',
                             'source', 'commit') if k in tj}
    of['file'] = 'profiles/' + cands[-1]
    return tj.get('dram_bytes_per_launch'), of


# --------------------------------------------------------------------------------------
# the other BASELINE configs
# --------------------------------------------------------------------------------------
def bench_forward_b32(dev, with_cpu):
    """config 1: RGBF_EmbeddingModel.embed on 32 frames, host ndarray in -> host ndarray out"""
    from vpd_b200 import RGBF_EmbeddingModel
    cpu, x = cpu_embed_fps() if with_cpu else (None, None)
    if x is None:
        x = torch.randn((32, 5, IMG, IMG), generator=torch.Generator().manual_seed(0)).numpy()
    torch.manual_seed(0)
    m = RGBF_EmbeddingModel('resnet34', EMB, True, 'cuda')
    for _ in range(5):
        m.embed(x)
    torch.cuda.synchronize()
    reps = 30
    t0 = time.perf_counter()
    for _ in range(reps):
        out = m.embed(x)
    dt = (time.perf_counter() - t0) / reps
    xd = torch.from_numpy(x).to(dev)
    m.eval()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m(xd)
    e1.record()
    torch.cuda.synchronize()
    dms = e0.elapsed_time(e1) / reps
    return {'what': 'BASELINE config 1: embed() of 32 frames, fp32 [32,5,128,128] host array in, '
                    'float32 [32,32] host array out (H2D + layout conversion + eval forward + D2H)',
            'value': 32 / dt, 'unit': UNIT, 'ms_per_call': dt * 1e3,
            'device_only': {'value': 32 / dms * 1e3, 'ms_per_call': dms,
                            'tflops': FWD_FLOP_PER_FRAME * 32 / dms / 1e9},
            'out_shape': list(out.shape), 'cpu_baseline': cpu}


def bench_keypoint(dev, with_cpu, pk):
    """config 4: Keypoint_EmbeddingModel.epoch at n = 4096 (see tests/diag_keypoint.py)"""
    from oracle import keypoint_train_ref as T          # synthetic batch generator + CPU leg only
    from vpd_b200 import keypoint
    from vpd_b200.keypoint_train import FCPoseDecoder
    N, HID, BLOCKS = 4096, 1024, 2
    step_gflop = 3 * (3 * 8.534 + 2 * 0.700) * N / 1e3   # fwd MFLOP/sample x3 (train)
    torch.manual_seed(0)
    enc = keypoint.FCResNet(39, 32, BLOCKS, HID, dropout=0.2)
    dec = FCPoseDecoder(32, [512, 512], [('h36m', 140)])
    cpu_enc = {k: v.clone() for k, v in enc.state_dict().items()}
    cpu_dec = {k: v.clone() for k, v in dec.state_dict().items()}
    model = keypoint.Keypoint_EmbeddingModel(enc, {'3d': dec}, 'cuda')
    opt = model.get_optimizer(1e-4)
    batches = [{k: v.to(dev) for k, v in T.synth_batch(N, 100 + i).items()} for i in range(4)]
    model.epoch([('h36m', batches[:3])], optimizer=opt)
    torch.cuda.synchronize()
    steps = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    _, loss, _ = model.epoch([('h36m', [batches[i % 4] for i in range(steps)])], optimizer=opt)
    e1.record()
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / steps
    res = {'what': 'BASELINE config 4: VIPE* keypoint-embedding training step (3 weight-sharing '
                   'encoder passes 39->1024x2 blocks->32, 3-D pose decoder, hinge + MSE, backward, '
                   'AdamW), n = 4096 synthetic poses, through Keypoint_EmbeddingModel.epoch',
           'value': N / ms * 1e3, 'unit': 'samples/s', 'ms_per_step': ms, 'loss': loss,
           'roofline': {'bound': 'tensor', 'achieved': round(step_gflop / ms, 2),
                        'peak': pk['tflops_burst'], 'unit': 'TFLOP/s',
                        'frac': round(step_gflop / ms / pk['tflops_burst'], 4),
                        'note': '81 GFLOP per step in 4096x1024x1024 GEMMs: launch- and '
                                'HBM-bound (25 MB activation tensors), not tensor-bound'}}
    if with_cpu:
        torch.set_num_threads(os.cpu_count())
        b = T.synth_batch(N, 100)
        masks = [T.replay_masks(N, HID, BLOCKS, 3, 0.2)]
        t0 = time.perf_counter()
        T.zipped_step(cpu_enc, cpu_dec, dec.fcn_keys, [('h36m', b)], masks, 0.2, BLOCKS)
        dt = time.perf_counter() - t0
        res['cpu_baseline'] = {'value': N / dt, 'unit': 'samples/s', 'cores': os.cpu_count(),
                               'kind': 'port', 'sample': '1 step of 4096 samples, forward + '
                               'backward (no optimizer), torch fp32, oracle port'}
    return res


def bench_corpus(enc, dev, rank, world, barrier, dist, apply_dev_fps):
    """config 5, scaled: every rank embeds ceil(371/8) = 47 synthetic videos of 2,695 frames
    (1 M frames over 8 GPUs) from pinned host memory and writes their pickles to local disk."""
    from vpd_b200 import apply as vapply, synth
    per_rank, frames = 47, 2695
    distinct = 4
    gen = torch.Generator().manual_seed(900 + rank)
    pool_rgb = [torch.randint(0, 256, (frames, IMG, IMG, 3), generator=gen, dtype=torch.uint8).pin_memory()
                for _ in range(distinct)]
    pool_flow = [torch.randint(0, 256, (frames, IMG, IMG, 3), generator=gen, dtype=torch.uint8).pin_memory()
                 for _ in range(distinct)]
    videos = []
    for r in range(world):
        for v in range(per_rank):
            # only this rank's videos carry data; the others are placeholders of the same length
            mine = r == rank
            videos.append(('r{}v{:03d}'.format(r, v), list(range(frames)),
                           pool_rgb[v % distinct] if mine else pool_rgb[0],
                           pool_flow[v % distinct] if mine else pool_flow[0]))
    out_dir = tempfile.mkdtemp(prefix='vpd_corpus_r{}_'.format(rank))
    try:
        warm = [('warm', list(range(1200)), pool_rgb[0][:1200], pool_flow[0][:1200])]
        vapply.extract_corpus(enc, warm, out_dir, synth.FS_MEAN_STD, flip=True, writers=0)
        barrier()
        timing = {}
        t0 = time.perf_counter()
        # equal-length videos: the longest-first partition gives every rank `per_rank` of them
        # (which ones does not matter - the names are unique, the pixels are this rank's pool)
        names = vapply.extract_corpus(enc, videos, out_dir, synth.FS_MEAN_STD, flip=True,
                                      rank=rank, world_size=world, writers=3, timing=timing)
        barrier()
        dt = time.perf_counter() - t0
        n_files = len([f for f in os.listdir(out_dir) if f.endswith('.emb.pkl')]) - 1
        nbytes = sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir))
        import pickle
        with open(os.path.join(out_dir, names[0] + '.emb.pkl'), 'rb') as fp:
            one = pickle.load(fp)
        ok = (len(one) == frames and one[0][0] == 0 and one[0][1].shape == (2, EMB)
              and str(one[0][1].dtype) == 'float32' and one[0][2] == {})
    finally:
        shutil.rmtree(out_dir, ignore_errors=True)
    if dist is not None:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    total = per_rank * frames * world
    fps = total / dt
    return {'what': 'BASELINE config 5 scaled to {} GPU(s): {} videos x {} frames per rank '
                    '({} distinct videos of random pixels per rank in pinned host memory, the '
                    'rest alias them), [orig, flipped] -> eval encoder -> per-video '
                    '(frame, float32[2,32], {{}}) pickles on local disk; wall clock around '
                    'extract_corpus, barrier on both sides, max over ranks'.format(
                        world, per_rank, frames, distinct),
            'value': fps, 'unit': UNIT, 'seconds': dt, 'frames': total,
            'videos_written_rank0': n_files, 'pickle_bytes_rank0': nbytes, 'pickle_ok': bool(ok),
            'h2d_bytes_per_frame': IMG * IMG * 6,
            'vs_device_only_apply': round(fps / apply_dev_fps, 3) if apply_dev_fps else None}


# --------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------
def run_native(args, rank, world, local_rank):
    from vpd_b200 import synth, RGBF_EmbeddingModel, ModelTrainer
    from vpd_b200._lib import lib
    from vpd_b200.assemble import assemble_stem, assemble_batch
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    dist = None
    numa_cpus = None
    if world > 1:
        import torch.distributed as dist
        from vpd_b200 import dp
        if os.environ.get('VPD_NUMA_BIND', '1') != '0':
            numa_cpus = dp.bind_host_to_device(local_rank)   # before any pinned allocation
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch

    torch.manual_seed(0)
    enc = RGBF_EmbeddingModel('resnet34', EMB, True, 'cuda')
    trainer = ModelTrainer(enc, True)
    opt, _ = trainer.get_optimizer(5e-4)

    # device-resident synthetic pools (rank-offset seeds), larger than L2
    rgb, flow = synth.crops(POOL, seed=1 + 100 * rank)
    teach = synth.teacher(POOL, seed=3 + 100 * rank, emb_dim=EMB, motion=True)
    rgb, flow, teach = rgb.to(dev), flow.to(dev), teach.to(dev)
    NIDX = 64
    gi = torch.Generator().manual_seed(4 + 100 * rank)
    idx_all = torch.randint(0, POOL, (NIDX, B), generator=gi).int().to(dev)
    flip_all = torch.randint(0, 2, (NIDX, B), generator=gi).to(torch.uint8).to(dev)
    tgt = torch.empty((B, 2 * EMB), device=dev)
    stem = trainer.stem_buffer(B, IMG, IMG)

    def step(i):
        j = i % NIDX
        assemble_stem(stem, rgb, flow, synth.FS_MEAN_STD, flip=flip_all[j], teacher=teach,
                      index=idx_all[j], tgt=tgt)
        trainer.train_step_stem(stem, tgt, B, IMG, IMG, opt)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    # preheat: the same steps for >= preheat_s seconds, so that the timed region runs at the
    # clocks / power state of a long training run (every rank runs the same count)
    n_pre = 0
    if args.preheat_s > 0:
        t0 = time.perf_counter()
        for i in range(8):
            step(args.warmup + i)
        torch.cuda.synchronize()
        per = (time.perf_counter() - t0) / 8
        n_pre = int(math.ceil(args.preheat_s / per))
        if dist is not None:
            t = torch.tensor([n_pre], device=dev, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_pre = int(t.item())
        for i in range(n_pre):
            step(args.warmup + 8 + i)
        n_pre += 8
    barrier()
    sampler.mark()
    L = lib()
    launches0 = L.call('vpd_launch_count')
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trainer._loss.zero_()
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + n_pre + i)
    ev1.record()
    barrier()
    launches = L.call('vpd_launch_count') - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps
    value = B * world * args.steps / (ms / 1e3)
    final_loss = trainer._loss.item() / (B * args.steps)

    # ---- data-parallel correctness of one extra step (every rank) -------------------------
    dp_check = None
    if dist is not None:
        d = assemble_batch(rgb, flow, synth.FS_MEAN_STD, flip=flip_all[0], teacher=teach,
                           index=idx_all[0])
        chk = trainer.dp_self_check(d['img'], d['emb'], B)
        flag = torch.tensor([0.0 if chk['ok'] else 1.0, chk['max_rel']], device=dev,
                            dtype=torch.float64)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        dp_check = {'ok': bool(flag[0].item() == 0.0), 'max_rel': flag[1].item(),
                    'buckets': chk['buckets'], 'overlapped': chk['overlapped'], 'rtol': 1e-4,
                    'what': 'per all-reduce bucket: |arena after the trainer step - sum over ranks '
                            'of the ranks\' own gradients| / |sum|, max over buckets and ranks'}
        del d
        barrier()

    # ---- kernel times (rank 0): CUPTI activity records of 3 more steps run like the timed ones,
    #      and CUDA events around every launch of an eager single-stream pass as a cross-check --
    roofline, breakdown = None, None
    pk = peaks()
    if rank == 0:
        net = enc._native(IMG, IMG, B)
        if dist is not None:       # rank-0-only passes: no collectives may be issued from them
            L.call('vpd_net_set_bucket_callback', net.handle, None, None)
            trainer._hooked = None
        nprof = 3

        def local_step(i):
            j = i % NIDX
            assemble_stem(stem, rgb, flow, synth.FS_MEAN_STD, flip=flip_all[j], teacher=teach,
                          index=idx_all[j], tgt=tgt)
            L.call('vpd_net_train_step', net.handle, None, stem, tgt, B, trainer._loss,
                   torch.cuda.current_stream().cuda_stream)
            opt.step()

        # one serialised pass (eager, single stream, an event pair around every launch) read
        # two ways: CUPTI kernel durations (kernel only) and the event intervals (kernel + gap)
        L.call('vpd_net_profile_enable', net.handle, 1)
        local_step(0)
        torch.cuda.synchronize()
        L.call('vpd_net_profile_read', net.handle, (ctypes.c_float * 64)(), (ctypes.c_int * 64)())
        L.call('vpd_net_profile_enable', net.handle, 1)
        cupti = cupti_kernel_times(local_step, nprof)
        torch.cuda.synchronize()
        msbuf = (ctypes.c_float * 64)()
        cntbuf = (ctypes.c_int * 64)()
        L.call('vpd_net_profile_read', net.handle, msbuf, cntbuf)
        L.call('vpd_net_profile_enable', net.handle, 0)
        ev_ms = [sum(msbuf[k * 8:(k + 1) * 8]) / nprof for k in range(8)]
        ev_n = [sum(cntbuf[k * 8:(k + 1) * 8]) // nprof for k in range(8)]
        f_fwd, f_dgrad, f_wgrad = conv_flops(B)
        flops = {'conv_fwd': f_fwd, 'conv_dgrad': f_dgrad, 'conv_wgrad': f_wgrad}
        n_params = enc._params.numel()
        fam_bytes = {}
        # the HBM-bound kernels north_star names, each timed alone (CUDA events around 20
        # back-to-back launches; inputs larger than L2 / the 600 MB optimizer state)
        # (replayed from a CUDA graph: the Python wrappers cost more host time per call than the
        # ~20 us kernels run, so an eager loop would time the launch rate, not the kernels)
        def loop_ms(fn, reps=20):
            for _ in range(3):
                fn(0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    for r in range(reps):
                        fn(r)
                graph.replay()
                torch.cuda.synchronize()
                e0.record()
                graph.replay()
                e1.record()
            except Exception:
                torch.cuda.synchronize()
                e0.record()
                for r in range(reps):
                    fn(r)
                e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        def hbm_entry(ms_, nbytes, what):
            gbs = nbytes / (ms_ * 1e-3) / 1e9
            return {'ms_per_launch': round(ms_, 4), 'what': what,
                    'roofline': {'bound': 'hbm', 'achieved': round(gbs, 1), 'peak': pk['hbm_gbs'],
                                 'unit': 'GB/s', 'frac': round(gbs / pk['hbm_gbs'], 4),
                                 'algorithmic_bytes_per_launch': int(nbytes)},
                    'frac': round(gbs / pk['hbm_gbs'], 4)}

        img_f32 = torch.empty((B, 1, 5, IMG, IMG), device=dev, dtype=torch.float32)
        alone = {
            'assemble_stem': hbm_entry(
                loop_ms(lambda r: assemble_stem(stem, rgb, flow, synth.FS_MEAN_STD,
                                                flip=flip_all[r % NIDX], teacher=teach,
                                                index=idx_all[r % NIDX], tgt=tgt)),
                K1_BYTES_PER_FRAME * B,
                'K1, training batch: 256 random frames of the uint8 pools -> bf16 network layout '
                '(81,920 B read + 291,584 B written per frame)'),
            'assemble_nchw': hbm_entry(
                loop_ms(lambda r: assemble_batch(rgb, flow, synth.FS_MEAN_STD,
                                                 flip=flip_all[r % NIDX], teacher=teach,
                                                 index=idx_all[r % NIDX])),
                409600 * B,
                'K1, reference layout: the same frames -> fp32 [B,5,128,128] (409,600 B per frame; '
                'includes the output allocation)'),
            'adamw': hbm_entry(loop_ms(lambda r: opt.step()), ADAMW_BYTES_PER_PARAM * n_params,
                               'K5 over the flat arena of {} parameters (28 B each)'.format(n_params)),
        }
        del img_f32
        breakdown = {}
        if cupti is not None:
            for name, d in sorted(cupti.items()):
                per_ms = d['us'] / nprof / 1e3
                entry = {'ms_per_step': round(per_ms, 4), 'launches': d['n'] // nprof,
                         'by_kernel_us_per_step': {k: [v[0] // nprof, round(v[1] / nprof, 1)]
                                                   for k, v in sorted(d['by_kernel'].items())}}
                if name in flops:
                    entry['tflops'] = round(flops[name] / (per_ms * 1e-3) / 1e12, 2)
                    entry['frac_burst'] = round(entry['tflops'] / pk['tflops_burst'], 4)
                if name in fam_bytes:
                    gbs = fam_bytes[name] / (per_ms * 1e-3) / 1e9
                    entry['roofline'] = {'bound': 'hbm', 'achieved': round(gbs, 1),
                                         'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                                         'frac': round(gbs / pk['hbm_gbs'], 4),
                                         'algorithmic_bytes_per_launch': fam_bytes[name]}
                    entry['frac'] = entry['roofline']['frac']
                breakdown[name] = entry
            breakdown['_how'] = ('CUPTI activity records (hardware start/end timestamps per launch) of '
                                 '{} steps run serialised after the timed region (eager, one stream: '
                                 'no programmatic early start, no side-stream overlap), caches in '
                                 'their in-step state'.format(nprof))
        breakdown['_alone'] = alone
        events = {}
        for k, name in enumerate(KINDS):
            if ev_n[k]:
                events[name] = {'ms_per_step': round(ev_ms[k], 4), 'launches': ev_n[k]}
        breakdown['_events'] = events
        if cupti is not None and 'conv_fwd' in cupti and 'conv_dgrad' in cupti:
            t_ig = (cupti['conv_fwd']['us'] + cupti['conv_dgrad']['us']) / nprof / 1e3
            n_ig = (cupti['conv_fwd']['n'] + cupti['conv_dgrad']['n']) // nprof
            how = breakdown['_how']
        else:
            t_ig, n_ig = ev_ms[0] + ev_ms[2], ev_n[0] + ev_n[2]
            how = ('CUDA events around every launch on the launching stream, {} eager '
                   'single-stream steps after the timed region (includes launch gaps)'.format(nprof))
        f_ig = f_fwd + f_dgrad
        achieved = f_ig / (t_ig * 1e-3) / 1e12
        achieved_ev = f_ig / ((ev_ms[0] + ev_ms[2]) * 1e-3) / 1e12
        # which measured peak applies: the run is pre-heated and the GPU sits at its power cap
        # (throttle reason sw_power_cap, ~990 W) for the whole timed region - the conditions the
        # driver's SUSTAINED figure was measured under (cuBLAS, 998 W; MEASURED_PEAKS.json); the
        # burst figure applies to a run at full clocks that never reaches the cap
        power_capped = bool(clocks and 'sw_power_cap' in (clocks.get('reasons') or []))
        full_clocks = bool(clocks and clocks.get('sm_mhz') and clocks.get('sm_max_mhz')
                           and clocks['sm_mhz'] >= 0.9 * clocks['sm_max_mhz']) and not power_capped
        peak = pk['tflops_burst'] if full_clocks else pk['tflops']
        traffic, traffic_of = latest_traffic()
        roofline = {'kernel': 'conv_igemm_kernel + conv3x3_halo_kernel + conv3x3_halo_stream_kernel '
                              '(all forward and data-gradient launches of the step)',
                    'bound': 'tensor', 'achieved': round(achieved, 2), 'peak': peak,
                    'unit': 'TFLOP/s', 'frac': round(achieved / peak, 4),
                    'frac_burst': round(achieved / pk['tflops_burst'], 4),
                    'frac_sustained': round(achieved / pk['tflops'], 4),
                    'peak_source': '{}: bf16 {} (timed region ran at {} of {} MHz, {})'.format(
                        pk['source'], 'burst' if full_clocks else 'sustained',
                        clocks.get('sm_mhz') if clocks else None,
                        clocks.get('sm_max_mhz') if clocks else None,
                        'power-capped at {} W'.format(clocks.get('power_w_max')) if power_capped
                        else 'not power-capped'),
                    'traffic': traffic, 'traffic_of': traffic_of,
                    'launches_per_step': n_ig, 'avg_launch_ms': round(t_ig / max(1, n_ig), 5),
                    'algorithmic_flops_per_step': f_ig,
                    'achieved_events': round(achieved_ev, 2), 'how': how}

    # ---- e2e through ModelTrainer.epoch with pinned HOST batches ------------------------
    # `e2e` = the reference loader's own fp32 {'img','emb'} batches (the drop-in number);
    # `e2e.u8_batches` = raw uint8 crops + flip bits + both teacher rows (K1 on the device)
    e2e = None
    if not args.no_e2e:
        if dist is not None:
            trainer._hooked = None
        nb = 3
        host_u8, host_f32 = [], []
        for j in range(nb):
            idx = idx_all[j].long()
            host_u8.append({'rgb_u8': rgb[idx].cpu().pin_memory(),
                            'flow_u8': flow[idx].cpu().pin_memory(),
                            'flip': flip_all[j].cpu().pin_memory(),
                            'teacher': teach[idx].cpu().pin_memory(),
                            'rgb_mean_std': synth.FS_MEAN_STD})
            b = {'img': torch.empty((B, 5, IMG, IMG), dtype=torch.float32).pin_memory(),
                 'emb': torch.empty((B, 2 * EMB), dtype=torch.float32).pin_memory()}
            d = assemble_batch(rgb, flow, synth.FS_MEAN_STD, flip=flip_all[j], teacher=teach,
                               index=idx_all[j])
            b['img'].copy_(d['img'])
            b['emb'].copy_(d['emb'])
            host_f32.append(b)
        torch.cuda.synchronize()
        e2e_steps = max(4, min(args.steps, 20))

        def timed_epoch(host):
            trainer.epoch([host[j % nb] for j in range(3)], optimizer=opt)      # warm-up
            barrier()
            t0 = time.perf_counter()
            ev0.record()
            loss = trainer.epoch([host[j % nb] for j in range(e2e_steps)], optimizer=opt)
            ev1.record()
            barrier()
            wall = time.perf_counter() - t0
            ms_e = max(ev0.elapsed_time(ev1), wall * 1e3)   # includes the loss read-back
            if dist is not None:
                t = torch.tensor([ms_e], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_e = t.item()
            return B * world * e2e_steps / (ms_e / 1e3), loss

        fps_f32, loss_f32 = timed_epoch(host_f32)
        fps_u8, loss_u8 = timed_epoch(host_u8)
        u8_bytes = sum(v.numel() * v.element_size() for k, v in host_u8[0].items()
                       if isinstance(v, torch.Tensor))
        e2e = {'value': fps_f32, 'unit': UNIT,
               'h2d_bytes_per_step': B * (5 * IMG * IMG + 2 * EMB) * 4, 'd2h_bytes_per_step': 8,
               'steps': e2e_steps,
               'api': "ModelTrainer.epoch(loader of pinned host fp32 {'img','emb'} batches as the "
                      "reference's DataLoader yields them, optimizer): H2D copy, layout "
                      "conversion, train step, AdamW, loss read-back",
               'loss_per_frame': loss_f32,
               'u8_batches': {'value': fps_u8, 'unit': UNIT, 'h2d_bytes_per_step': u8_bytes,
                              'api': "ModelTrainer.epoch(loader of pinned host batches {'rgb_u8',"
                                     "'flow_u8','flip','teacher'}, optimizer): H2D copy, K1 "
                                     "assembly, train step, AdamW, loss read-back",
                              'loss_per_frame': loss_u8}}
        del host_u8, host_f32

    # ---- apply path (apply_vpd_model.py): K1 [orig, flipped] -> eval-mode encoder ----
    apply_res = None
    if not args.no_e2e:
        from vpd_b200 import apply as vapply
        enc.eval()
        nfr = vapply.BATCH_SIZE
        net_a = enc._native(IMG, IMG, 2 * nfr)
        stem_a = L.call('vpd_net_stem_input', net_a.handle)

        def apply_step(i):
            lo = (i * nfr) % (POOL - nfr)
            assemble_stem(stem_a, rgb[lo:lo + nfr], flow[lo:lo + nfr], synth.FS_MEAN_STD, k=2)
            return enc.embed_stem(stem_a, 2 * nfr, IMG, IMG)

        for i in range(3):
            apply_step(i)
        barrier()
        n_apply = 10
        ev0.record()
        for i in range(n_apply):
            apply_step(3 + i)
        ev1.record()
        barrier()
        ms_a = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms_a], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_a = t.item()
        apply_res = {'value': nfr * world * n_apply / (ms_a / 1e3), 'unit': 'frames/s',
                     'images_per_s': 2 * nfr * world * n_apply / (ms_a / 1e3),
                     'batch_frames': nfr, 'variants_per_frame': 2,
                     'tflops_per_gpu': round(2 * FWD_FLOP_PER_FRAME * nfr * n_apply / (ms_a / 1e3) / 1e12, 1),
                     'what': 'device-timed: uint8 crops in HBM -> [orig, flipped] assembly -> '
                             'eval-mode ResNet-34 encoder -> fp32 [2,32] embeddings in HBM'}

    # ---- the other BASELINE configs -------------------------------------------------------
    configs = None
    if not args.no_configs:
        configs = {}
        with_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
        try:
            configs['corpus_apply'] = bench_corpus(enc, dev, rank, world, barrier, dist,
                                                   apply_res['value'] if apply_res else None)
        except Exception as exc:          # keep the headline line even if a side config breaks
            configs['corpus_apply'] = {'error': '{}: {}'.format(type(exc).__name__, exc)}
            barrier()
        if rank == 0:
            for key, fn in (('forward_b32', lambda: bench_forward_b32(dev, with_cpu)),
                            ('keypoint_train_n4096', lambda: bench_keypoint(dev, with_cpu, pk))):
                try:
                    configs[key] = fn()
                except Exception as exc:
                    configs[key] = {'error': '{}: {}'.format(type(exc).__name__, exc)}
    enc.train()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_train_fps(steps=2, warmup=1, batch=B)
        cpu = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}

    if rank == 0:
        step_tflops = TRAIN_FLOP_PER_FRAME * value / world / 1e12
        cfg = workload_config(world, B)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': cfg, 'preheat_steps': n_pre, 'clocks': clocks, 'e2e': e2e,
            'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
            'dp_check': dp_check, 'kernels': breakdown, 'apply': apply_res, 'configs': configs,
            'step_tflops_per_gpu': round(step_tflops, 2),
            'step_frac_of_bf16_burst_peak': round(step_tflops / pk['tflops_burst'], 4),
            'loss_per_frame_timed_region': final_loss,
            'host_numa_bind': ({'rank0_cpus': len(numa_cpus), 'how': 'vpd_b200.dp.bind_host_to_device: every '
                                'rank pinned to the CPUs of its GPU\'s NUMA node before allocating pinned batches'}
                               if numa_cpus else None),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = get_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
               '--nproc-per-node', str(args.gpus), '--master-addr', '127.0.0.1',
               '--master-port', os.environ.get('MASTER_PORT', '29533')] + sys.argv
        sys.exit(subprocess.call(cmd))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the native arm has no CPU fallback)')
    run_native(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
