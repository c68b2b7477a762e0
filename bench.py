#!/usr/bin/env python3
"""Benchmark of the VPD student hot path (BASELINE.json metric: student train frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one distillation training step of the ResNet-34 student on a batch of
256 synthetic 128x128 RGB+flow crops per GPU (BASELINE.json configs[1]): K1 batch
assembly from device-resident uint8 pools -> train-mode forward -> sum-MSE loss ->
backward -> [NCCL all-reduce SUM of the gradient arena when N > 1] -> fused AdamW.

  value   device-timed frames/s over all ranks (CUDA events, max over ranks), inputs
          already in HBM;
  e2e     the same metric through the reference-facing call `ModelTrainer.epoch`
          fed with HOST (pinned) batches of raw uint8 crops (what the path's first stage,
          K1, consumes), H2D copies, K1 and the loss read-back inside the timed region;
          e2e.fp32_batches = the same with the fp32 {'img','emb'} batches the reference's
          DataLoader yields (4x the bytes over PCIe);
  roofline  the dominant kernel family, measured live with CUDA events around every
          launch in a separate profiled pass of the same step;
  cpu_baseline  the oracle port of the reference's fp32 PyTorch path on the host cores
          (rank 0, N = 1 only, bounded sample).
`--impl reference` times only that CPU path (the reference is pure PyTorch; its own
CPU implementation of this path is what oracle/student_ref.py restates).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
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
METRIC = 'vpd_student_train_frames_per_s'
UNIT = 'frames/s'


def get_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', type=str, default='native', choices=['native', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fp:
            p = json.load(fp)
        return {'hbm_gbs': p['hbm_gbs'], 'tflops': p['bf16_tflops_sustained'],
                'tflops_burst': p['bf16_tflops'], 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops': 1400.0, 'tflops_burst': 1590.0, 'source': 'fallback'}


# --------------------------------------------------------------------------------------
# CPU arm: the reference's own fp32 PyTorch path, restated in oracle/student_ref.py
# --------------------------------------------------------------------------------------
def cpu_train_fps(steps, warmup, budget_s, batch=BATCH):
    """frames/s of OracleTrainer.step on the host cores; the per-step sample batch is
    shrunk (never below 8) so that steps+warmup fit in `budget_s`."""
    from oracle import assemble_ref, student_ref
    from vpd_b200 import synth
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = student_ref.init_encoder_state('resnet34', EMB, True)
    dsd = student_ref.init_decoder_state(EMB)
    tr = student_ref.OracleTrainer(sd, dsd, lr=5e-4)
    rgb, flow = synth.crops(batch, seed=1)
    teach = synth.teacher(batch, seed=3, emb_dim=EMB, motion=True)
    fl = synth.flips(batch, seed=2)
    img, tgt = assemble_ref.train_batch(rgb.numpy(), flow.numpy(), teach.numpy(), fl.numpy(),
                                        *synth.FS_MEAN_STD)
    # probe to size the sample
    t0 = time.perf_counter()
    tr.step(img[:16], tgt[:16])
    per_frame = (time.perf_counter() - t0) / 16
    sample = int(budget_s / max(1, steps + warmup) / per_frame)
    sample = max(8, min(batch, sample - sample % 8))
    for _ in range(warmup):
        tr.step(img[:sample], tgt[:sample])
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.step(img[:sample], tgt[:sample])
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {'value': sample * steps / total, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '{} train steps of {} frames (fp32, torch CPU, {} threads; oracle port of '
                      'ModelTrainer.epoch)'.format(steps, sample, cores),
            'ms_per_step': 1e3 * total / steps, 'sample_batch': sample}


def run_reference(args, rank):
    if rank != 0:
        return
    res = cpu_train_fps(args.steps, max(1, min(args.warmup, 2)), budget_s=150.0)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': res['value'], 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': workload_config(1, res['sample_batch']),
        'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': res['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
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

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = self.rows[getattr(self, 'first', 0):]
        if len(rows) < 2:
            rows = self.rows[-4:]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# --------------------------------------------------------------------------------------
# native arm
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


def run_native(args, rank, world, local_rank):
    from vpd_b200 import synth, RGBF_EmbeddingModel, ModelTrainer
    from vpd_b200._lib import lib
    from vpd_b200.assemble import assemble_stem
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch

    torch.manual_seed(0)                      # identical init on every rank
    enc = RGBF_EmbeddingModel('resnet34', EMB, True, 'cuda')
    trainer = ModelTrainer(enc, True)
    opt, _ = trainer.get_optimizer(5e-4)

    # device-resident synthetic pools (rank-offset seeds), larger than L2
    rgb, flow = synth.crops(POOL, seed=1 + 100 * rank)
    teach = synth.teacher(POOL, seed=3 + 100 * rank, emb_dim=EMB, motion=True)
    rgb, flow, teach = rgb.to(dev), flow.to(dev), teach.to(dev)
    total_steps = args.warmup + args.steps + 8
    gi = torch.Generator().manual_seed(4 + 100 * rank)
    idx_all = torch.randint(0, POOL, (total_steps, B), generator=gi).int().to(dev)
    flip_all = torch.randint(0, 2, (total_steps, B), generator=gi).to(torch.uint8).to(dev)
    tgt = torch.empty((B, 2 * EMB), device=dev)
    stem = trainer.stem_buffer(B, IMG, IMG)

    def step(i):
        assemble_stem(stem, rgb, flow, synth.FS_MEAN_STD, flip=flip_all[i], teacher=teach,
                      index=idx_all[i], tgt=tgt)
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
    sampler.mark()
    L = lib()
    launches0 = L.call('vpd_launch_count')
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trainer._loss.zero_()
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
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

    # ---- profiled pass: CUDA events around every kernel of the step (rank 0 only) ----
    roofline, breakdown = None, None
    pk = peaks()
    if rank == 0:
        net = enc._native(IMG, IMG, B)
        nprof = 3
        # rank-0-only pass: no collectives may be issued from it
        L.call('vpd_net_set_bucket_callback', net.handle, None, None)
        trainer._hooked = None
        L.call('vpd_net_profile_enable', net.handle, 1)
        for i in range(nprof):
            # same body as step() but without the collective, so only kernels are timed
            assemble_stem(stem, rgb, flow, synth.FS_MEAN_STD, flip=flip_all[i], teacher=teach,
                          index=idx_all[i], tgt=tgt)
            L.call('vpd_net_train_step', net.handle, None, stem, tgt, B, trainer._loss,
                   torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        msbuf = (ctypes.c_float * 64)()
        cntbuf = (ctypes.c_int * 64)()
        L.call('vpd_net_profile_read', net.handle, msbuf, cntbuf)
        L.call('vpd_net_profile_enable', net.handle, 0)
        per_kind_ms = [sum(msbuf[k * 8:(k + 1) * 8]) / nprof for k in range(8)]
        per_kind_n = [sum(cntbuf[k * 8:(k + 1) * 8]) // nprof for k in range(8)]
        f_fwd, f_dgrad, f_wgrad = conv_flops(B)
        flops = {0: f_fwd, 2: f_dgrad, 3: f_wgrad}
        breakdown = {}
        for k, name in enumerate(KINDS):
            if per_kind_n[k] == 0:
                continue
            entry = {'ms_per_step': round(per_kind_ms[k], 4), 'launches': per_kind_n[k]}
            if k in flops and per_kind_ms[k] > 0:
                entry['tflops'] = round(flops[k] / (per_kind_ms[k] * 1e-3) / 1e12, 2)
            stages = {}
            for s in range(8):
                if cntbuf[k * 8 + s]:
                    stages[str(s)] = round(msbuf[k * 8 + s] / nprof, 4)
            entry['by_stage_ms'] = stages
            breakdown[name] = entry
        # dominant kernel = conv_igemm_kernel / conv3x3_halo_kernel, the tensor-core implicit
        # GEMM that runs every forward convolution and every data gradient (the dgrad
        # launches also carry the fused BN-backward reduction); wgrad is listed beside it
        t_ig = per_kind_ms[0] + per_kind_ms[2]
        f_ig = flops[0] + flops[2]
        n_ig = per_kind_n[0] + per_kind_n[2]
        achieved = f_ig / (t_ig * 1e-3) / 1e12
        # DRAM traffic per launch of the family's largest member, from the committed ncu
        # --set full capture (profiles/): not measurable live without the profiler
        traffic, traffic_of = None, None
        tpath = os.path.join(ROOT, 'profiles', 'r01d_traffic.json')
        if os.path.exists(tpath):
            with open(tpath) as fp:
                tj = json.load(fp)
            traffic = tj['dram_bytes_per_launch']
            traffic_of = {k: tj[k] for k in ('kernel', 'algorithmic_bytes_per_launch',
                                             'tensor_pipe_active_pct', 'source')}
        roofline = {'kernel': 'conv_igemm_kernel + conv3x3_halo_kernel (forward and dgrad launches)',
                    'bound': 'tensor', 'achieved': round(achieved, 2), 'peak': pk['tflops'],
                    'unit': 'TFLOP/s', 'frac': round(achieved / pk['tflops'], 4),
                    'traffic': traffic, 'traffic_of': traffic_of,
                    'peak_source': pk['source'] + ' bf16 sustained',
                    'launches_per_step': n_ig,
                    'avg_launch_ms': round(t_ig / max(1, n_ig), 5),
                    'algorithmic_flops_per_step': f_ig,
                    'forward_only_tflops': round(flops[0] / (per_kind_ms[0] * 1e-3) / 1e12, 2),
                    'wgrad_kernel_tflops': round(flops[3] / (per_kind_ms[3] * 1e-3) / 1e12, 2),
                    'how': 'CUDA events around every launch on the launching stream, '
                           '{} profiled steps after the timed region'.format(nprof)}

    # ---- e2e through ModelTrainer.epoch with pinned HOST batches ------------------------
    # (a) raw uint8 crops + flip bits + both teacher rows: what the path's first stage (K1)
    #     consumes; normalise / stack / flip / row select run on the device  -> `e2e`
    # (b) the reference loader's own fp32 {'img', 'emb'} batches (4x the bytes) -> e2e.fp32_batches
    e2e = None
    if not args.no_e2e:
        from vpd_b200.assemble import assemble_batch
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
        e2e = {'value': fps_u8, 'unit': UNIT, 'h2d_bytes_per_step': u8_bytes,
               'd2h_bytes_per_step': 8, 'steps': e2e_steps,
               'api': "ModelTrainer.epoch(loader of pinned host batches {'rgb_u8','flow_u8',"
                      "'flip','teacher'}, optimizer): H2D copy, K1 assembly, train step, "
                      "AdamW, loss read-back",
               'loss_per_frame': loss_u8,
               'fp32_batches': {'value': fps_f32, 'unit': UNIT,
                                'h2d_bytes_per_step': B * (5 * IMG * IMG + 2 * EMB) * 4,
                                'api': "ModelTrainer.epoch(loader of pinned host fp32 "
                                       "{'img','emb'} batches as the reference's DataLoader "
                                       "yields them, optimizer)",
                                'loss_per_frame': loss_f32}}

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
            out_a = apply_step(3 + i)
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
                     'what': 'device-timed: uint8 crops in HBM -> [orig, flipped] assembly -> '
                             'eval-mode ResNet-34 encoder -> fp32 [2,32] embeddings in HBM'}
        enc.train()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_train_fps(steps=2, warmup=1, budget_s=25.0)
        cpu = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}

    if rank == 0:
        step_tflops = TRAIN_FLOP_PER_FRAME * value / world / 1e12
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': workload_config(world, B), 'clocks': clocks, 'e2e': e2e,
            'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
            'kernels': breakdown, 'apply': apply_res,
            'step_tflops_per_gpu': round(step_tflops, 2),
            'step_frac_of_bf16_peak': round(step_tflops / pk['tflops'], 4),
            'loss_per_frame_timed_region': final_loss,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
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
