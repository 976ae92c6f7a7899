import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.lib()
    return orc


def pytest_sessionfinish(session, exitstatus):
    """RCB_DEBUG_SO=1: the suite ran on the -DRCB_BOUNDS build; any index that left its array fails the session and the
    verdict is written where the GPU box's caller collects it (gpurun_out/bounds_report.json)."""
    if os.environ.get("RCB_DEBUG_SO") != "1":
        return
    import json

    try:
        import recast_rs_b200.builder as rb

        ctx = rb.Context(0)
        rep = ctx.bounds_report()
        ctx.close()
    except Exception as e:  # no GPU here: nothing ran
        rep = {"instrumented": None, "error": str(e)}
    rep["tests_exit_status"] = int(exitstatus)
    rep["tests_collected"] = int(getattr(session, "testscollected", 0))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "bounds_report.json"), "w"))
    print("\n[rcb bounds] " + json.dumps(rep))
    if rep.get("violations"):
        session.exitstatus = 1
