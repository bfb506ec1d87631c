"""Which shape of the CPU pool gives the reference arm its best whole-pool rate on this host (config 2, bounded
sample of bench.CpuArm)?  Host-only; run on the GPU box to measure the cores the bench's CPU arm runs on.
    timeout 300 python scripts/cpu_pool.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
    cfg = bench.CONFIGS['c2']
    for procs in (64, 32, 16, 8):
        arm = bench.CpuArm(cfg, procs, 'P1')
        try:
            fps, wall, est, t_f, t_x = arm.step()
            print(f"procs={procs} blas_threads={arm.blas_threads}: {fps:.3f} frames/s (wall {wall:.1f}s, est {est:.1f} s/frame/worker, "
                  f"LS sample {t_f:.1f}s, xambg {t_x:.1f}s)", flush=True)
        finally:
            arm.close()


if __name__ == "__main__":
    main()
