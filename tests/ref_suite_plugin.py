"""pytest plugin used by tests/test_plugin_reference.py::test_reference_own_suite_over_host_loops (build container
only).  Loaded with `-p tests.ref_suite_plugin` into a pytest run of the REFERENCE'S OWN test suite
(/root/reference/tests): at session start it calls cola_b200.install(), routes CPU operators through the adapters
(FORCE_FAST_PATH) and enters tests/host_harness.emulated_kernels(), so every torch float32/float64 case of the
reference's tests that reaches a Krylov loop or a structured matmat runs on this package's host orchestration
(over the test-only kernel statements).  At session end it writes how often each kernel wrapper was called to
$COLA_B200_REF_SUITE_STATS, which is how the caller knows the fast path was really taken."""
import collections
import json
import os

_STATE = {}


def pytest_sessionstart(session):
    import cola

    import cola_b200.backend as be
    from cola_b200 import plugin
    from tests.host_harness import _WRAPPERS, emulated_kernels
    plugin.install(cola)
    plugin.FORCE_FAST_PATH = True
    cm = emulated_kernels()
    cm.__enter__()
    calls = collections.Counter()

    def counted(name, fn):
        def wrapper(*a, **k):
            calls[name] += 1
            return fn(*a, **k)
        return wrapper

    for name in _WRAPPERS:
        setattr(be, name, counted(name, getattr(be, name)))
    lib = be.lib()
    lib.call = counted("cola_cg_*", lib.call)
    _STATE.update(cm=cm, calls=calls, plugin=plugin)


def pytest_sessionfinish(session, exitstatus):
    if not _STATE:
        return
    path = os.environ.get("COLA_B200_REF_SUITE_STATS")
    if path:
        with open(path, "w") as fh:
            json.dump(dict(_STATE["calls"]), fh)
    _STATE["cm"].__exit__(None, None, None)
    _STATE["plugin"].FORCE_FAST_PATH = False
    _STATE["plugin"].uninstall()
