"""CPU tests of the drop-in boundary: libctcb.so loads without a GPU, exports every symbol
include/*.h declares, sizes workspaces, validates arguments, and REFUSES to compute without a
CUDA device (there is no CPU fallback on the product path)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build_cuda()
    from gluon_e2e_asr_b200 import _lib
    return _lib


def _declared():
    names = set()
    for h in ("ctcb.h", "ctcb_dlpack.h"):
        with open(os.path.join(ROOT, "include", h)) as f:
            src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
        names |= set(re.findall(r"\b(ctcb_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported(lib):
    l = lib.load()
    names = _declared()
    assert len(names) >= 13
    for n in sorted(names):
        assert hasattr(l, n), "libctcb.so does not export %s" % n
    # and the Python binding's own list is the header's list
    assert set(lib.EXPORTS) == names


def test_version_and_error_string(lib):
    l = lib.load()
    assert l.ctcb_version() == 103
    assert isinstance(l.ctcb_last_error(), bytes)


def test_problem_struct_layout_matches_header(lib):
    """ctypes mirror of ctcb_problem_t: field order/size as the C compiler lays it out."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "ctcb.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(ctcb_problem_t), offsetof(ctcb_problem_t, logits), offsetof(ctcb_problem_t, grad),' \
          'offsetof(ctcb_problem_t, labels), offsetof(ctcb_problem_t, head_grad), offsetof(ctcb_problem_t, status),' \
          'offsetof(ctcb_problem_t, logits_row_offsets));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "l.c")
        with open(c, "w") as f:
            f.write(src)
        exe = os.path.join(td, "l")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    P = lib.Problem
    assert got == [ctypes.sizeof(P), P.logits.offset, P.grad.offset, P.labels.offset, P.head_grad.offset,
                   P.status.offset, P.logits_row_offsets.offset]


def test_proj_struct_layout_matches_header(lib):
    """ctypes mirror of ctcb_proj_t (the output projection in front of the loss)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "ctcb.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(ctcb_proj_t), offsetof(ctcb_proj_t, hidden_stride_b), offsetof(ctcb_proj_t, K),' \
          'offsetof(ctcb_proj_t, weight), offsetof(ctcb_proj_t, bias), offsetof(ctcb_proj_t, operand_dtype));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "l.c")
        with open(c, "w") as f:
            f.write(src)
        exe = os.path.join(td, "l")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    P = lib.Proj
    assert got == [ctypes.sizeof(P), P.hidden_stride_b.offset, P.K.offset, P.weight.offset, P.bias.offset, P.operand_dtype.offset]


def test_proj_entry_validation_without_gpu(lib):
    """ctcb_proj_forward refuses NULL arguments before it touches a device."""
    l = lib.load()
    p = lib.Problem()
    assert l.ctcb_proj_forward(None, ctypes.byref(p), 0, None, 0, None) == lib.CTCB_INVALID_VALUE
    assert b"proj" in l.ctcb_last_error()
    assert l.ctcb_proj_loss_grad(None, ctypes.byref(p), None, 0, None) == lib.CTCB_INVALID_VALUE


def test_workspace_bytes(lib):
    a = lib.workspace_bytes(500, 32, 46, 120, True)
    b = lib.workspace_bytes(500, 32, 46, 120, False)
    c = lib.workspace_bytes(500, 64, 46, 120, True)
    assert 0 < b < a < c
    assert a % 256 == 0
    with pytest.raises(lib.CtcbError) as e:
        lib.workspace_bytes(0, 32, 46, 120)
    assert e.value.code == lib.CTCB_INVALID_VALUE


def test_argument_validation_without_gpu(lib):
    l = lib.load()
    p = lib.Problem()
    assert l.ctcb_loss_grad(ctypes.byref(p), None, 0, None) == lib.CTCB_INVALID_VALUE
    assert b"bad shape" in l.ctcb_last_error()
    x = np.zeros((2, 3, 4), np.float32)
    loss = np.zeros((2,), np.float32)
    p.T, p.B, p.V, p.Lmax, p.blank = 3, 2, 4, 0, 7
    p.logits, p.loss = x.ctypes.data, loss.ctypes.data
    assert l.ctcb_loss_grad(ctypes.byref(p), None, 0, None) == lib.CTCB_INVALID_VALUE
    assert b"blank" in l.ctcb_last_error()
    p.blank = 0
    p.Lmax, p.labels = 5000, x.ctypes.data
    assert l.ctcb_loss_grad(ctypes.byref(p), None, 0, None) == lib.CTCB_UNSUPPORTED


def test_no_cpu_fallback(lib):
    """Host pointers on the device entry are an error, and so is a CPU tensor at the Python
    surface; the host-buffer entry needs a CUDA device."""
    torch = pytest.importorskip("torch")
    from gluon_e2e_asr_b200 import CtcLoss, ctc_loss
    with pytest.raises(RuntimeError, match="no CPU path"):
        ctc_loss(torch.zeros(3, 2, 4), torch.zeros(2, 1))
    with pytest.raises(RuntimeError, match="no CPU path"):
        CtcLoss()(torch.zeros(2, 3, 4), torch.zeros(2, 1))
    if torch.cuda.is_available():
        pytest.skip("the rest checks behaviour on a box without a GPU")
    l = lib.load()
    x = np.zeros((2, 3, 4), np.float32)
    loss = np.zeros((2,), np.float32)
    p = lib.Problem()
    p.T, p.B, p.V, p.Lmax = 3, 2, 4, 0
    p.logits, p.logits_stride_t, p.logits_stride_b = x.ctypes.data, 4, 12
    p.loss = loss.ctypes.data
    ws = np.zeros((1 << 16,), np.uint8)
    addr = (ws.ctypes.data + 255) // 256 * 256
    assert l.ctcb_loss_grad(ctypes.byref(p), addr, 1 << 15, None) != lib.CTCB_OK
    assert l.ctcb_loss_grad_host(ctypes.byref(p), 0) in (lib.CTCB_UNSUPPORTED, lib.CTCB_EXECUTION_FAILED)
    assert np.all(loss == 0)


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "gluon_e2e_asr_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dp, f)) as fh:
                    s = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), f
                assert "ctc_ref" not in s, f


def test_block_constructor_contract():
    """loss.py:111-119: layout assertions and batch axis."""
    pytest.importorskip("torch")
    from gluon_e2e_asr_b200 import CtcLoss
    assert CtcLoss(layout="NTC", label_layout="NT")._batch_axis == 0
    assert CtcLoss(layout="TNC", label_layout="TN")._batch_axis == 1
    with pytest.raises(AssertionError):
        CtcLoss(layout="NCT")
    with pytest.raises(AssertionError):
        CtcLoss(label_layout="NN")
    with pytest.raises(ValueError):
        CtcLoss(blank_label="middle")


def test_pipe_entry_validation(lib):
    """ctcb_pipe_*: argument errors are reported, and without a CUDA device the pipe cannot be made
    (no CPU path behind the prefetching host entry either)."""
    torch = pytest.importorskip("torch")
    l = lib.load()
    h = ctypes.c_void_p()
    assert l.ctcb_pipe_create(0, 2, None) == lib.CTCB_INVALID_VALUE
    assert l.ctcb_pipe_create(0, 0, ctypes.byref(h)) == lib.CTCB_INVALID_VALUE
    assert l.ctcb_pipe_create(0, 9, ctypes.byref(h)) == lib.CTCB_INVALID_VALUE
    t = ctypes.c_int64(0)
    assert l.ctcb_pipe_submit(None, None, ctypes.byref(t)) == lib.CTCB_INVALID_VALUE
    assert l.ctcb_pipe_wait(None, 0, None) == lib.CTCB_INVALID_VALUE
    assert l.ctcb_pipe_destroy(None) == lib.CTCB_OK
    if not torch.cuda.is_available():
        assert l.ctcb_pipe_create(0, 2, ctypes.byref(h)) == lib.CTCB_UNSUPPORTED
        assert not h.value
        from gluon_e2e_asr_b200 import HostPipeline
        with pytest.raises(RuntimeError):
            HostPipeline(0)
