import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Set in the child process that runs ONE `unverified` test (see _run_isolated): the child runs the test body itself.
_CHILD_ENV = "GB_UNVERIFIED_CHILD"
UNVERIFIED_TIMEOUT_S = int(os.environ.get("GB_UNVERIFIED_TIMEOUT", 300))
_HUNG = []  # an isolated test that hung: the remaining ones are skipped (bounded cost of never-run code: one timeout)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "unverified: GPU test of code that has not run on a B200 yet; collected LAST, run "
                                       "in its own process under a timeout, outcome recorded as xfail / xpass so "
                                       "that neither a failure nor a hang of never-run code takes the verified suite "
                                       "down (GB_UNVERIFIED_STRICT=1: ordinary pass / fail)")
    config.addinivalue_line("markers", "experimental: GPU test of an opt-in kernel path that has never run on a B200; "
                                       "skipped unless GB_EXPERIMENTAL=1")


def _keep_log(item, text):
    """Failure output of an isolated test, kept where a GPU visit's artefacts are collected (gpurun_out/)."""
    try:
        d = os.path.join(ROOT, "gpurun_out", "unverified")
        os.makedirs(d, exist_ok=True)
        name = "".join(ch if ch.isalnum() or ch in "-_." else "_" for ch in item.name)
        with open(os.path.join(d, name + ".log"), "w") as f:
            f.write(text)
    except OSError:
        pass


def _run_isolated(item):
    """Run one test in a child pytest process: a kernel of never-run code that spins on an mbarrier forever (or
    faults the context) costs that child, not the process that holds the rest of the suite's CUDA context."""
    def runtest():
        if _HUNG:
            pytest.skip(f"not run: {_HUNG[0]} hung earlier in this session")
        env = dict(os.environ, **{_CHILD_ENV: "1", "GB_UNVERIFIED_STRICT": "1"})
        cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
               f"{item.fspath}::{item.name}"]
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=UNVERIFIED_TIMEOUT_S, env=env, cwd=ROOT)
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
            _keep_log(item, f"hung: killed after {UNVERIFIED_TIMEOUT_S} s\n{out}")
            _HUNG.append(item.name)
            pytest.fail(f"hung: killed after {UNVERIFIED_TIMEOUT_S} s\n{out[-3000:]}", pytrace=False)
        if res.returncode != 0:
            _keep_log(item, res.stdout + "\n" + res.stderr)
            pytest.fail(f"child pytest exit code {res.returncode}\n{res.stdout[-6000:]}\n{res.stderr[-2000:]}",
                        pytrace=False)
    return runtest


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda it: 1 if ("unverified" in it.keywords or "experimental" in it.keywords) else 0)  # stable
    if os.environ.get("GB_EXPERIMENTAL", "0") != "1":
        skip_exp = pytest.mark.skip(reason="opt-in kernel path not yet verified on a B200 (GB_EXPERIMENTAL=1 runs it)")
        for item in items:
            if "experimental" in item.keywords:
                item.add_marker(skip_exp)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu or os.environ.get("GB_UNVERIFIED_FORCE_ISOLATION") == "1":
        strict = os.environ.get("GB_UNVERIFIED_STRICT", "0") == "1"
        for item in items:
            if "unverified" not in item.keywords:
                continue
            if os.environ.get(_CHILD_ENV) != "1":
                item.runtest = _run_isolated(item)
            if not strict:
                item.add_marker(pytest.mark.xfail(strict=False, reason="first B200 run of this code: outcome recorded "
                                                                       "(x = failed / hung, X = passed)"))
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
