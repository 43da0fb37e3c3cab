"""Drop-in aliasing: make ``third_party.a2c_ppo_acktr.{model,storage,distributions,utils,algo,algo.ppo,
algo.gail,baselines.common.running_mean_std}`` resolve to this package.

The reference's caller (third_party/a2c_ppo_acktr/main_gail_dyn_ppo.py:32-38) imports the hot-path
classes by those module paths, and its checkpoints are whole-object pickles that record the class path
``third_party.a2c_ppo_acktr.model.Policy`` (main_gail_dyn_ppo.py:307-320, my_pybullet_envs/utils.py:24-56).
``install()`` registers alias modules in ``sys.modules`` so that both keep working unmodified; modules
that are NOT on the hot path (``arguments``, ``envs``, the rest of ``baselines``) still
come from the reference tree when ``reference_root`` is given.
"""
import importlib
import os
import sys
import types

_A2C = "third_party.a2c_ppo_acktr"
ALIASES = {
    _A2C + ".model": "simgan_b200.model",
    _A2C + ".model_split": "simgan_b200.model_split",
    _A2C + ".storage": "simgan_b200.storage",
    _A2C + ".distributions": "simgan_b200.distributions",
    _A2C + ".utils": "simgan_b200.utils",
    _A2C + ".algo": "simgan_b200.algo",
    _A2C + ".algo.ppo": "simgan_b200.algo.ppo",
    _A2C + ".algo.gail": "simgan_b200.algo.gail",
    _A2C + ".baselines.common.running_mean_std": "simgan_b200.running_mean_std",
}


def _package(name, path):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    if path and os.path.isdir(path) and path not in mod.__path__:
        mod.__path__.append(path)
    return mod


def install(reference_root=None):
    """Register the alias modules.  ``reference_root``: checkout of the reference whose non-hot-path
    modules should stay importable under the same package names (optional)."""
    ref = reference_root
    pk = {
        "third_party": ref and os.path.join(ref, "third_party"),
        _A2C: ref and os.path.join(ref, "third_party", "a2c_ppo_acktr"),
        _A2C + ".baselines": ref and os.path.join(ref, "third_party", "a2c_ppo_acktr", "baselines"),
        _A2C + ".baselines.common": ref and os.path.join(ref, "third_party", "a2c_ppo_acktr", "baselines", "common"),
    }
    for name, path in pk.items():
        _package(name, path)
    for alias, target in ALIASES.items():
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        parent, _, leaf = alias.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    return sorted(ALIASES)


def uninstall():
    for alias in ALIASES:
        sys.modules.pop(alias, None)
