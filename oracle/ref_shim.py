"""Import the UNMODIFIED reference modules from /root/reference (dev container only).

TEST INFRASTRUCTURE. Used only by ``oracle/make_golden.py`` and by the
``-m "not gpu"`` tests that pin the oracle restatement against the real
reference when ``/root/reference`` is mounted (it is not on the GPU box; every
use is guarded by :func:`available`).

The reference's hot-path modules import cleanly once three non-path modules are
stubbed:

* ``third_party.a2c_ppo_acktr.envs``   (pulls in ``gym``; only ``VecNormalize``
  is looked up, by ``third_party/a2c_ppo_acktr/utils.py:29``)
* ``pybullet``                          (``my_pybullet_envs/utils.py:20``)
* package ``my_pybullet_envs``          (its ``__init__`` registers gym envs,
  ``my_pybullet_envs/__init__.py:15-22``); ``utils.py`` is loaded by path.

No reference file is modified or copied.
"""
import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SIMGAN_REFERENCE_ROOT", "/root/reference")
_A2C = "third_party.a2c_ppo_acktr"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "third_party", "a2c_ppo_acktr", "storage.py"))


class _RefModules(types.SimpleNamespace):
    pass


_cache = None


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Return a namespace with the reference's storage/model/distributions/ppo/gail/rms modules.

    The reference package is imported under the private prefix ``_simgan_ref`` so that it never
    collides with this repo's own ``third_party.a2c_ppo_acktr`` alias package.  Because the
    reference uses absolute imports (``from third_party.a2c_ppo_acktr.utils import ...``) the
    canonical names are bound temporarily while importing and restored afterwards.
    """
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REF_ROOT)

    saved = {k: v for k, v in sys.modules.items()
             if k == "third_party" or k.startswith("third_party.") or k in ("pybullet", "my_pybullet_envs")
             or k.startswith("my_pybullet_envs.")}
    for k in saved:
        del sys.modules[k]
    saved_path = list(sys.path)
    try:
        sys.path.insert(0, REF_ROOT)
        # stub 1: envs (gym-dependent) -> only VecNormalize is referenced
        envs_stub = types.ModuleType(_A2C + ".envs")
        envs_stub.VecNormalize = type("VecNormalize", (), {})
        # stub 2: pybullet
        pb_stub = types.ModuleType("pybullet")
        # stub 3: bare package my_pybullet_envs + utils loaded by path
        pkg = types.ModuleType("my_pybullet_envs")
        pkg.__path__ = [os.path.join(REF_ROOT, "my_pybullet_envs")]
        sys.modules["pybullet"] = pb_stub
        sys.modules["my_pybullet_envs"] = pkg
        importlib.import_module("third_party")
        importlib.import_module(_A2C)
        sys.modules[_A2C + ".envs"] = envs_stub
        env_utils = _load_by_path("my_pybullet_envs.utils", os.path.join(REF_ROOT, "my_pybullet_envs", "utils.py"))
        ns = _RefModules(
            storage=importlib.import_module(_A2C + ".storage"),
            utils=importlib.import_module(_A2C + ".utils"),
            distributions=importlib.import_module(_A2C + ".distributions"),
            model=importlib.import_module(_A2C + ".model"),
            model_split=importlib.import_module(_A2C + ".model_split"),
            ppo=importlib.import_module(_A2C + ".algo.ppo"),
            gail=importlib.import_module(_A2C + ".algo.gail"),
            rms=importlib.import_module(_A2C + ".baselines.common.running_mean_std"),
            env_utils=env_utils,
        )
        ns.modules = {k: v for k, v in sys.modules.items()
                      if k == "third_party" or k.startswith("third_party.") or k == "pybullet"
                      or k == "my_pybullet_envs" or k.startswith("my_pybullet_envs.")}
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k == "third_party" or k.startswith("third_party.") or k in ("pybullet", "my_pybullet_envs") \
                    or k.startswith("my_pybullet_envs."):
                del sys.modules[k]
        sys.modules.update(saved)
    _cache = ns
    return ns


class BoxSpace:
    """Minimal stand-in for gym.spaces.Box; the reference only reads ``__class__.__name__`` and
    ``.shape`` (third_party/a2c_ppo_acktr/model.py:56-58, storage.py:41-44)."""

    def __init__(self, dim):
        self.shape = (int(dim),)


BoxSpace.__name__ = "Box"
