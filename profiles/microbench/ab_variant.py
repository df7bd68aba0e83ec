"""A/B a kernel variant against the default build on the GPU box.

Here (no GPU):   python profiles/microbench/ab_variant.py ILB_RASTER_BALLOT ILB_FAST_UNORM8
  builds illuminant_b200/libilluminant_b200_variant.so with those defines (the .so is git-ignored but travels with gpurun)
  and prints the gpurun command that runs the parity tests and the N2 / N3 microbenchmark with both libraries.
The library a process loads is chosen by ILB_LIB (illuminant_b200/_abi.py); ILB_OUT / ILB_DEFINES steer illuminant_b200/build.py."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
VARIANT = ROOT / "illuminant_b200" / "libilluminant_b200_variant.so"


def main():
    defines = sys.argv[1:]
    if not defines:
        sys.exit(__doc__)
    env = dict(os.environ, ILB_DEFINES=" ".join(defines), ILB_OUT=str(VARIANT))
    subprocess.run([sys.executable, "-c", "from illuminant_b200 import build; print(build.build(force=True))"], cwd=ROOT, env=env, check=True)
    rel = VARIANT.relative_to(ROOT)
    tests = "tests/test_particle_render.py tests/test_gpu_resolve.py"
    print("\n/usr/local/graft/bin/gpurun --timeout 240 -- '"
          f"ILB_LIB=$PWD/{rel} python -m pytest {tests} -m gpu -q 2>&1 | tail -3; "
          "python profiles/microbench/n2n3_profile.py > gpurun_out/ab_default.json; "
          f"ILB_LIB=$PWD/{rel} python profiles/microbench/n2n3_profile.py > gpurun_out/ab_variant.json; "
          "cat gpurun_out/ab_default.json gpurun_out/ab_variant.json'")


if __name__ == "__main__":
    main()
