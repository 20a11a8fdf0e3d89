"""The UNMODIFIED reference modules as a checker / baseline (TEST AND BASELINE INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.py and bench.py's CPU legs may import this file; nothing under gtos_b200/ does.

The reference (jcyk/gtos) is pure Python.  `build_ref()` is the recipe that stages it for runs on the GPU box, where
/root/reference does not exist: it copies the reference's own files, byte for byte, from where they lie into
`oracle/_ref/{generator,translator}/` - a git-ignored OUTPUT directory (like a compiled oracle/_ref/*.so would be; it
is never committed, but it travels with the gpurun snapshot).  `load()` imports them from there (or straight from
/root/reference when the copy is absent) under private module names, optionally with `gtos_b200/dropin` in front so
that the reference's generator.py assembles its Generator out of the B200 modules exactly as INTEGRATION.md describes.

Run-time shims for torch 2.x (the files themselves are not edited, no numeric effect):
  * transformer.MultiheadAttention.in_proj_qkv returns clones (`q *= scaling` is in-place on a chunk view,
    transformer.py:120)
  * transformer.SelfAttentionMask.forward returns a bool mask (the reference builds uint8, transformer.py:212, which
    torch 2.x masked_fill_ rejects)
  * `cpu_cuda_noop()`: Tensor.cuda is a no-op while the reference's beam search runs on the CPU (search.py:72,136 and
    generator.py:117 call `.cuda(device)` unconditionally)
"""
import contextlib
import hashlib
import importlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference"
REF_OUT = os.path.join(HERE, "_ref")
DROPIN = os.path.join(ROOT, "gtos_b200", "dropin")
TASKS = ("generator", "translator")
# the model + search + host-side helpers generator.py imports; nothing else of the reference is staged
FILES = ("graph_transformer.py", "transformer.py", "encoder.py", "decoder.py", "generator.py", "search.py", "data.py",
         "utils.py", "adam.py")
_FLAT = ("generator", "encoder", "decoder", "transformer", "graph_transformer", "data", "search", "utils", "adam",
         "_gtos_reference_encoder")


def build_ref(verbose=False):
    """stage the reference's files under oracle/_ref/ (needs /root/reference: the build container).  Returns the
    staged directory, or None when the reference tree is not present (GPU box: the staged copy travelled)."""
    if not os.path.isdir(REF_SRC):
        return REF_OUT if have_ref() else None
    for task in TASKS:
        dst = os.path.join(REF_OUT, task)
        os.makedirs(dst, exist_ok=True)
        for f in FILES:
            src = os.path.join(REF_SRC, task, f)
            if os.path.exists(src):
                shutil.copyfile(src, os.path.join(dst, f))
    with open(os.path.join(REF_OUT, "MANIFEST.txt"), "w") as fh:
        fh.write("staged by oracle/ref_loader.py::build_ref from /root/reference (unmodified copies; not committed)\n")
        for task in TASKS:
            for f in FILES:
                p = os.path.join(REF_OUT, task, f)
                if os.path.exists(p):
                    fh.write(f"{task}/{f} sha256 {hashlib.sha256(open(p, 'rb').read()).hexdigest()}\n")
    if verbose:
        print(open(os.path.join(REF_OUT, "MANIFEST.txt")).read())
    return REF_OUT


def ref_dir(task="generator"):
    staged = os.path.join(REF_OUT, task)
    if os.path.exists(os.path.join(staged, "generator.py")):
        return staged
    live = os.path.join(REF_SRC, task)
    if os.path.exists(os.path.join(live, "generator.py")):
        return live
    return None


def have_ref(task="generator"):
    return ref_dir(task) is not None


_cache = {}


def load(task="generator", dropin=False):
    """-> namespace(generator, search, data, transformer, graph_transformer, encoder, decoder) of the reference's
    modules for `task`; dropin=True: the reference's generator.py / search.py / data.py over the gtos_b200 modules."""
    key = (task, bool(dropin))
    if key in _cache:
        return _cache[key]
    d = ref_dir(task)
    if d is None:
        raise FileNotFoundError("reference modules not found: run oracle/ref_loader.py::build_ref() in the build "
                                "container (oracle/_ref/) or provide /root/reference")
    saved = {n: sys.modules.pop(n) for n in _FLAT if n in sys.modules}
    old_path = list(sys.path)
    try:
        sys.path[:] = ([DROPIN, ROOT] if dropin else []) + [d] + [p for p in old_path if os.path.abspath(p or ".") != d]
        importlib.invalidate_caches()
        gen = importlib.import_module("generator")
        ns = types.SimpleNamespace(task=task, dropin=bool(dropin), dir=d, generator=gen,
                                   **{n: sys.modules[n] for n in ("search", "data", "transformer", "graph_transformer",
                                                                  "encoder", "decoder")})
        ns.adam = importlib.import_module("adam")
    finally:
        sys.path[:] = old_path
        for n in _FLAT:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    if not dropin:
        tf = ns.transformer
        orig_qkv = tf.MultiheadAttention.in_proj_qkv
        tf.MultiheadAttention.in_proj_qkv = lambda self, q: tuple(t.clone() for t in orig_qkv(self, q))
        orig_mask = tf.SelfAttentionMask.forward
        tf.SelfAttentionMask.forward = lambda self, size: orig_mask(self, size).bool()
    _cache[key] = ns
    return ns


@contextlib.contextmanager
def cpu_cuda_noop():
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


if __name__ == "__main__":
    print(build_ref(verbose=True))
