"""Times the UNMODIFIED reference (baseline/_ref, or /root/reference in the build container) on the host cores for
bench.py --impl reference: its own InferenceCore.step on the bench workload -- same weights (name-seeded synthetic), same
synthetic video, the same pre-filled steady-state bank injected through the reference's own KeyValueMemoryStore.add --
in a CHILD process (the reference's package is also called `cutie`).  Prints one JSON object on the last stdout line:
{"per_frame_s": [...], "threads": T, "root": "..."}.  Never imported by the product.

    python baseline/ref_bench_child.py '<json job>'      job: {H, W, K, mem_frames, top_k, steps, warmup, max_seconds, threads}
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import warnings
    warnings.filterwarnings('ignore')
    import torch
    from cutie_b200.utils.synth import synthetic_video          # data generation only
    from oracle import ref_harness as rh
    job = json.loads(sys.argv[1])
    torch.set_num_threads(int(job['threads']))
    ref = rh.load_reference()
    cfg = rh.reference_cfg(mem_every=5, max_mem_frames=job['mem_frames'], use_long_term=False, top_k=job['top_k'])
    net = rh.build_reference_model(cfg)
    K = job['K']
    objs = list(range(1, K + 1))
    frames, mask = synthetic_video(job['warmup'] + job['steps'] + 2, job['H'], job['W'], K, seed=0)
    proc = ref.InferenceCore(net, cfg=cfg)
    t_begin = time.perf_counter()
    with torch.inference_mode():
        proc.step(frames[0], mask, objects=objs)
        HW = (job['H'] // 16) * (-(-job['W'] // 16))
        total = (job['mem_frames'] - 2) * HW
        g = torch.Generator().manual_seed(1234)
        done = 0
        while done < total:                                       # same chunks as bench.synthetic_bank_chunks
            n = min(16 * HW, total - done)
            key, shr = torch.randn(1, 64, n, generator=g), 1 + torch.randn(1, 1, n, generator=g) ** 2
            vals = torch.randn(1, K, 256, n, generator=g)
            proc.memory.work_mem.add(key, {o: vals[:, i] for i, o in enumerate(objs)}, shr, None, as_permanent='no')
            done += n
        print(f'[ref] bank prefilled: {proc.memory.work_mem.size(0)} tokens; threads={torch.get_num_threads()}',
              file=sys.stderr, flush=True)
        per_frame, t = [], 1
        for i in range(job['warmup'] + job['steps']):
            if per_frame and (time.perf_counter() - t_begin) + max(per_frame) > job['max_seconds'] and i >= job['warmup'] + 1:
                break
            t0 = time.perf_counter()
            proc.step(frames[t])
            t += 1
            dt = time.perf_counter() - t0
            if i >= job['warmup']:
                per_frame.append(dt)
            print(f'[ref] frame {i} {"(warmup) " if i < job["warmup"] else ""}{dt:.2f} s', file=sys.stderr, flush=True)
    print(json.dumps({'per_frame_s': per_frame, 'threads': torch.get_num_threads(), 'root': rh.REF_ROOT}))


if __name__ == '__main__':
    main()
