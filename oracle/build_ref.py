"""Recipe for oracle/_ref: a snapshot of the UNMODIFIED reference Python files the path needs, so that the reference itself
(not only the oracle port) can be run where /root/reference is not mounted (the GPU box): `bench.py --impl reference`
then times the real reference modules (`cpu_baseline.kind = "reference"`).

    python -m oracle.build_ref        # copies from /root/reference into oracle/_ref/ (git-ignored, travels with gpurun)

TEST / MEASUREMENT INFRASTRUCTURE ONLY. oracle/_ref/ is listed in .gitignore (reference sources never enter the history) and
nothing in the product path imports it; oracle/refimport.py falls back to it when /root/reference is absent.
Files (all read-only inputs of oracle/refimport.py's import recipe):
  CLAP/src/laion_clap/clap_module/{htsat,model,utils,feature_fusion,...}.py + model_configs/HTSAT-{tiny,base}.json
  CLAP/src/laion_clap/training/data.py     (get_audio_features / get_mel / int16 helpers are AST-lifted out of it)
  src/residual.py                          (ResiDual, patch_block_with_residual, setup_residual_htsat)
"""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "CLAP", "src", "laion_clap", "clap_module")):
        if verbose:
            print("oracle/build_ref: /root/reference is not mounted; keeping the existing oracle/_ref (if any)")
        return os.path.isdir(DST)
    rels = [os.path.relpath(p, SRC) for p in glob.glob(os.path.join(SRC, "CLAP", "src", "laion_clap", "clap_module", "*.py"))]
    rels += [os.path.join("CLAP", "src", "laion_clap", "clap_module", "model_configs", f"HTSAT-{n}.json") for n in ("tiny", "base")]
    rels += [os.path.join("CLAP", "src", "laion_clap", "training", "data.py"), os.path.join("src", "residual.py")]
    n = 0
    for r in rels:
        dst = os.path.join(DST, r)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, r), dst)
        n += 1
    if verbose:
        print(f"oracle/build_ref: {n} reference files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
