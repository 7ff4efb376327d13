"""TEST INFRASTRUCTURE: builds a ``DetectorW6`` launch plan without a GPU.  ``libb200track.so`` is replaced by a recorder whose
entry points return 0, tensors live on the CPU, and every ``b2t_conv_desc`` the planner would hand to the library is captured
(tests/test_detector_plan.py compares it with a committed golden plan and with the oracle's tensor shapes).  Nothing here
computes anything; the product never imports this module."""
import ctypes as C
from unittest import mock

import torch


class RecorderLib:
    def __init__(self):
        self.calls = []
        self.descs = []

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append(name)
            if name == "b2t_conv_plan_create":
                d = args[0]._obj
                self.descs.append({f: getattr(d, f) for f, _ in type(d)._fields_})
            if name == "b2t_conv_plan_flops":
                return 1.0
            if name == "b2t_nms_workspace_bytes":
                return 1 << 20
            if name.endswith("last_error"):
                return b""
            return 0
        return fn


def dry_run_plan(batch, img_size, **kw):
    """Returns (detector, [per-conv dict]) for a plan built on CPU tensors.  Pointers are reported relative to their tensors."""
    from b200track import _lib as L
    from b200track import detector as D
    from b200track.w6 import seeded_state_dict
    rec = RecorderLib()
    with mock.patch.object(L, "load", lambda: rec), mock.patch.object(torch.cuda, "is_available", lambda: True):
        det = D.DetectorW6(seeded_state_dict(0), batch=batch, img_size=img_size, device="cpu", use_graph=False, autotune=False, **kw)
    plans = [p for p in det.keep if hasattr(p, "keep") and isinstance(getattr(p, "keep"), tuple)]
    assert len(plans) == len(rec.descs)
    out = []
    for plan, d in zip(plans, rec.descs):
        x, w, b, y = plan.keep
        e = {k: v for k, v in d.items() if k not in ("x", "w_packed", "bias", "y")}
        e["x_off"] = int(d["x"]) - x.data_ptr()
        e["y_off"] = int(d["y"]) - y.data_ptr()
        e["x_shape"], e["y_shape"], e["w_shape"] = list(x.shape), list(y.shape), list(w.shape)
        e["y_dtype"] = str(y.dtype)
        out.append(e)
    return det, out
