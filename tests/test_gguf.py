"""GGUF -> device loader.  CPU: the container parser (pb200_gguf_probe) against files written by the upstream gguf-py writer
and against malformed files.  -m gpu: a model loaded by pb200_model_load_gguf decodes bit-identically to the same bytes handed over
tensor by tensor with pb200_model_set_tensor, whole and as a layer-window shard."""
import ctypes as C

import numpy as np
import pytest

from tiny_model import TinyModel


def _tiny(arch="llama", **kw):
    return TinyModel(n_layer=2, n_embd=512, n_head=4, n_head_kv=2, n_ff=1024, n_vocab=320, n_ctx=64, arch=arch,
                     ftype="q4_K_M" if arch == "llama" else "q5_K_M", seed=4, branch_scale=0.1, freq_factors=(arch == "llama"), **kw)


@pytest.mark.parametrize("arch", ["llama", "qwen2"])
def test_probe_reads_what_gguf_py_wrote(pkg, tmp_path, arch):
    tm = _tiny(arch)
    path = tmp_path / "m.gguf"
    tm.write_gguf(path)
    lib = pkg.Lib.get()
    hp = pkg.HParams()
    n, nbytes, a = C.c_int32(), C.c_int64(), C.create_string_buffer(16)
    assert lib.c.pb200_gguf_probe(str(path).encode(), C.byref(hp), C.byref(n), C.byref(nbytes), a) == 0
    assert a.value.decode() == arch
    assert n.value == len(tm.tensors)
    for k in ("n_layer", "n_embd", "n_head", "n_head_kv", "n_ff", "n_vocab", "rope_mode"):
        assert getattr(hp, k) == tm.hp[k], k
    assert hp.head_dim == 128 and hp.n_ctx_orig == tm.hp["n_ctx_orig"]
    assert abs(hp.rope_freq_base - tm.hp["rope_freq_base"]) < 1e-3 and abs(hp.rms_eps - tm.hp["rms_eps"]) < 1e-12
    # the data section holds at least every tensor's bytes (plus alignment padding)
    assert nbytes.value >= sum(np.asarray(a_).nbytes for _, a_ in tm.tensors.values())


def test_probe_rejects_bad_files(pkg, tmp_path):
    lib = pkg.Lib.get()
    hp = pkg.HParams()
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(b"GGML" + b"\0" * 64)
    assert lib.c.pb200_gguf_probe(str(bad).encode(), C.byref(hp), None, None, None) != 0        # wrong magic
    assert lib.c.pb200_gguf_probe(str(tmp_path / "missing.gguf").encode(), C.byref(hp), None, None, None) != 0
    good = tmp_path / "m.gguf"
    _tiny().write_gguf(good)
    raw = good.read_bytes()
    (tmp_path / "trunc.gguf").write_bytes(raw[: len(raw) // 2])                                  # tensor data cut off
    assert lib.c.pb200_gguf_probe(str(tmp_path / "trunc.gguf").encode(), C.byref(hp), None, None, None) != 0
    (tmp_path / "v1.gguf").write_bytes(raw[:4] + (1).to_bytes(4, "little") + raw[8:])           # version 1: not supported
    assert lib.c.pb200_gguf_probe(str(tmp_path / "v1.gguf").encode(), C.byref(hp), None, None, None) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["llama", "qwen2"])
def test_gguf_loaded_model_decodes_like_set_tensor_model(cuda, pkg, tmp_path, arch):
    tm = _tiny(arch)
    path = tmp_path / "m.gguf"
    tm.write_gguf(path)
    toks = [(i * 7919 + 13) % 320 for i in range(8)]
    ref = tm.load_engine(pkg)
    want = np.zeros((len(toks), 320), np.float32)
    for i, t in enumerate(toks):
        ref.decode(int(t), i, want[i])
    ref.close()
    eng = pkg.Model.from_gguf(path, n_ctx=64)
    assert eng.load_bytes == sum(np.asarray(a).nbytes for _, a in tm.tensors.values())
    got = np.zeros_like(want)
    for i, t in enumerate(toks):
        eng.decode(int(t), i, got[i])
    eng.close()
    assert np.array_equal(got, want)
    # a two-stage pipeline straight from the file: every shard reads only its window
    s0 = pkg.Model.from_gguf(path, layers=(0, 1), n_ctx=64)
    s1 = pkg.Model.from_gguf(path, layers=(1, 2), n_ctx=64)
    assert s0.load_bytes < eng.load_bytes and s1.load_bytes < eng.load_bytes
    got2 = np.zeros_like(want)
    for i, t in enumerate(toks):
        s0.decode(int(t), i, None)
        s1.set_hidden(s0.hidden())
        s1.decode(int(t), i, got2[i])
    s0.close(); s1.close()
    assert np.array_equal(got2, want)
