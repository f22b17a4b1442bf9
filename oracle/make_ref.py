"""Recipe for ``baseline/_ref/``: the REFERENCE itself, made runnable on the GPU box (test infrastructure, not product).

``/root/reference`` exists only in the authoring container. ``baseline/_ref/`` is git-ignored (no reference source ever
enters the history) but NOT gpurun-ignored, so whatever this script puts there travels to the B200 box with the
snapshot, like the built ``.so`` files do. It is used by exactly two things:

* ``bench.py --impl reference`` / the ``cpu_baseline`` leg: the reference's own PyTorch modules
  (``MSMFormer/meanshiftformer/modeling/**``, imported by path under ``tests/golden/ref_shim.py``) timed on the box's
  host cores - ``cpu_baseline.kind = "reference"`` instead of the slower oracle port;
* ``tools/bench_msda_ref.py``: the reference's own CUDA op (``pixel_decoder/ops/src``), built here for sm_100a, timed
  beside ``msm_ms_deform_attn_fused_fwd`` on the same box ("the kernel to beat").

What it does (idempotent; ``python oracle/make_ref.py [--force]``; ``__graft_entry__.build()`` calls it when
``/root/reference`` is present):

1. copies the ``*.py`` files of ``MSMFormer/meanshiftformer`` (and ``lib/networks``, ``lib/utils/mean_shift.py``)
   from ``/root/reference`` to ``baseline/_ref/reference/`` with the same relative paths;
2. copies ``pixel_decoder/ops/src`` to ``baseline/_ref/msda_src`` and applies the two-token patch the survey found
   necessary for torch >= 2.x: ``AT_DISPATCH_FLOATING_TYPES(value.type(), ...)`` -> ``value.scalar_type()``
   (``ms_deform_attn_cuda.cu:69,139``) - nothing else is touched, the kernels are the reference's;
3. builds it with ``torch.utils.cpp_extension.load`` for ``sm_100a`` (nvcc cross-compiles without a GPU, ~1 min) into
   ``baseline/_ref/msda_build/MultiScaleDeformableAttention.so``.
"""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MSM_REFERENCE_ROOT_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
DST_PY = os.path.join(DST, "reference")
DST_SRC = os.path.join(DST, "msda_src")
DST_BUILD = os.path.join(DST, "msda_build")
MSDA_SO = os.path.join(DST_BUILD, "MultiScaleDeformableAttention.so")
OPS_SRC = "MSMFormer/meanshiftformer/modeling/pixel_decoder/ops/src"


def copy_python():
    n = 0
    for sub in ("MSMFormer/meanshiftformer", "lib/networks"):
        for src in glob.glob(os.path.join(REF, sub, "**", "*.py"), recursive=True):
            rel = os.path.relpath(src, REF)
            dst = os.path.join(DST_PY, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            n += 1
    for rel in ("lib/utils/mean_shift.py",):
        dst = os.path.join(DST_PY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        n += 1
    return n


def _make_writable(top):
    for d, _, files in os.walk(top):
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)


def copy_and_patch_msda():
    if os.path.isdir(DST_SRC):
        _make_writable(DST_SRC)
        shutil.rmtree(DST_SRC)
    shutil.copytree(os.path.join(REF, OPS_SRC), DST_SRC)
    _make_writable(DST_SRC)  # /root/reference is mounted read-only and copytree keeps the modes
    cu = os.path.join(DST_SRC, "cuda", "ms_deform_attn_cuda.cu")
    text = open(cu).read()
    patched = text.replace("AT_DISPATCH_FLOATING_TYPES(value.type(),", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type(),")
    assert text.count("AT_DISPATCH_FLOATING_TYPES(value.type(),") == 2, "the reference changed: re-check the patch"
    open(cu, "w").write(patched)


def build_msda():
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    os.makedirs(DST_BUILD, exist_ok=True)
    srcs = ([os.path.join(DST_SRC, "vision.cpp")] + glob.glob(os.path.join(DST_SRC, "cpu", "*.cpp"))
            + glob.glob(os.path.join(DST_SRC, "cuda", "*.cu")))
    # the reference's own flags (ops/setup.py:44-49) + the sm_100a target; is_python_module=False: only build
    load(name="MultiScaleDeformableAttention", sources=srcs, extra_include_paths=[DST_SRC],
         extra_cflags=["-DWITH_CUDA"],
         extra_cuda_cflags=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                            "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                            "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=DST_BUILD, with_cuda=True, is_python_module=False, verbose=False)
    assert os.path.exists(MSDA_SO), MSDA_SO


def main(force=False):
    if not os.path.isdir(REF):
        print(f"make_ref: {REF} not present (GPU box): using the prebuilt baseline/_ref as is")
        return
    stamp = os.path.join(DST, ".stamp")
    if not force and os.path.exists(stamp) and os.path.exists(MSDA_SO):
        return
    n = copy_python()
    copy_and_patch_msda()
    build_msda()
    open(stamp, "w").write("reference d1c8487 vendored for the GPU box; see oracle/make_ref.py\n")
    print(f"make_ref: {n} python files -> {DST_PY}; MSDA op (patched 2 tokens) built for sm_100a -> {MSDA_SO}")


if __name__ == "__main__":
    main(force="--force" in sys.argv)
