"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference classes from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to generate tests/golden/*.pt fixtures and by `bench.py --impl reference`
when baseline/_ref or /root/reference is present.

Compat shim: the reference pins transformers<4.46 but the image has 5.5.0.
 (1) mantis/models/mllava/__init__.py imports processing_llava which imports removed hub helpers
     -> load configuration_llava.py / modeling_llava.py by file path under a synthetic package.
 (2) `_supports_sdpa` is a property reading self.language_model -> class attribute.
 (3) tie_weights() must accept kwargs.
 (4) shim subclasses must live in a real .py file (this one).
"""
import importlib.util
import os
import sys
import types

REF_ROOTS = [os.environ.get("MANTIS_REF_ROOT", ""), "/root/reference",
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "baseline", "_ref")]


def find_ref_root():
    for r in REF_ROOTS:
        if r and os.path.isfile(os.path.join(r, "mantis", "models", "mllava", "modeling_llava.py")):
            return os.path.abspath(r)
    return None


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def load_reference_mllava():
    """Returns (config_module, modeling_module) of the reference mllava, bypassing its __init__."""
    if "mllava" in _CACHE:
        return _CACHE["mllava"]
    root = find_ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (only available in the build container)")
    base = os.path.join(root, "mantis", "models", "mllava")
    pkg_name = "_mantis_ref.models.mllava"
    for p in ("_mantis_ref", "_mantis_ref.models", pkg_name):
        if p not in sys.modules:
            m = types.ModuleType(p)
            m.__path__ = []
            sys.modules[p] = m
    cfg = _load(pkg_name + ".configuration_llava", os.path.join(base, "configuration_llava.py"))
    mod = _load(pkg_name + ".modeling_llava", os.path.join(base, "modeling_llava.py"))
    _CACHE["mllava"] = (cfg, mod)
    return cfg, mod


def load_reference_idefics2():
    if "idefics2" in _CACHE:
        return _CACHE["idefics2"]
    root = find_ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (only available in the build container)")
    pkg_name = "_mantis_ref.models.idefics2"
    for p in ("_mantis_ref", "_mantis_ref.models", pkg_name):
        if p not in sys.modules:
            m = types.ModuleType(p)
            m.__path__ = []
            sys.modules[p] = m
    mod = _load(pkg_name + ".modeling_idefics2",
                os.path.join(root, "mantis", "models", "idefics2", "modeling_idefics2.py"))
    _CACHE["idefics2"] = mod
    return mod


def load_reference_idefics3():
    """-> (configuration module, modeling module) of mantis/models/idefics3, loaded by file path"""
    if "idefics3" in _CACHE:
        return _CACHE["idefics3"]
    root = find_ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (only available in the build container)")
    pkg_name = "_mantis_ref.models.idefics3"
    for p in ("_mantis_ref", "_mantis_ref.models", pkg_name):
        if p not in sys.modules:
            m = types.ModuleType(p)
            m.__path__ = []
            sys.modules[p] = m
    base = os.path.join(root, "mantis", "models", "idefics3")
    cfg = _load(pkg_name + ".configuration_idefics3", os.path.join(base, "configuration_idefics3.py"))
    mod = _load(pkg_name + ".modeling_idefics3", os.path.join(base, "modeling_idefics3.py"))
    _CACHE["idefics3"] = (cfg, mod)
    return cfg, mod


def ref_idefics3_classes():
    cfg, modm = load_reference_idefics3()

    class RefIdefics3(modm.Idefics3ForConditionalGeneration):
        _supports_sdpa = True

        def tie_weights(self, *a, **k):
            return None

    return cfg.Idefics3Config, cfg.Idefics3VisionConfig, RefIdefics3


def load_reference_llava_next():
    if "llava_next" in _CACHE:
        return _CACHE["llava_next"]
    root = find_ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (only available in the build container)")
    pkg_name = "_mantis_ref.models.mllava_next"
    for p in ("_mantis_ref", "_mantis_ref.models", pkg_name):
        if p not in sys.modules:
            m = types.ModuleType(p)
            m.__path__ = []
            sys.modules[p] = m
    mod = _load(pkg_name + ".modeling_llava_next",
                os.path.join(root, "mantis", "models", "mllava_next", "modeling_llava_next.py"))
    _CACHE["llava_next"] = mod
    return mod


def ref_llava_next_classes():
    modm = load_reference_llava_next()

    class RefLlavaNext(modm.LlavaNextForConditionalGeneration):
        _supports_sdpa = True

        def tie_weights(self, *a, **k):
            return None

    return modm.LlavaNextConfig, RefLlavaNext


def load_reference_processor():
    """mantis/models/mllava/processing_llava.py loaded by path.  It imports two hub helpers that transformers 5 removed
    (`is_remote_url`, `download_url`; only used by its from_pretrained override), so inert stand-ins are installed first."""
    if "processor" in _CACHE:
        return _CACHE["processor"]
    root = find_ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (only available in the build container)")
    import transformers
    import transformers.processing_utils as pu
    import transformers.utils.hub as hub
    for name in ("is_remote_url", "download_url", "is_offline_mode", "cached_file"):
        if not hasattr(hub, name):
            setattr(hub, name, lambda *a, **k: None)
    if not hasattr(pu, "transformers_module"):
        pu.transformers_module = transformers
    load_reference_mllava()                                    # creates the synthetic parent packages
    mod = _load("_mantis_ref.models.mllava.processing_llava",
                os.path.join(root, "mantis", "models", "mllava", "processing_llava.py"))
    _CACHE["processor"] = mod
    return mod


def load_reference_chat_utils():
    """mantis/models/mllava/utils.py (chat_mllava) with its relative imports satisfied by the modules loaded above."""
    if "chat_utils" in _CACHE:
        return _CACHE["chat_utils"]
    root = find_ref_root()
    load_reference_processor()
    if "_mantis_ref.models.conversation" not in sys.modules:
        _load("_mantis_ref.models.conversation", os.path.join(root, "mantis", "models", "conversation.py"))
    mod = _load("_mantis_ref.models.mllava.utils", os.path.join(root, "mantis", "models", "mllava", "utils.py"))
    _CACHE["chat_utils"] = mod
    return mod


def load_reference_train_data():
    """mantis/train/data.py (Collator, PackingDataset, ...) loaded by path.  Its video dependencies (av, decord) are absent
    from the image and unused by the classes the tests need -> empty stand-in modules; its `mantis.train.*` /
    `mantis.models.conversation` imports are satisfied with the reference's own files (the repo-root `mantis` package is the
    alias shim of the product and does not define them)."""
    if "train_data" in _CACHE:
        return _CACHE["train_data"]
    root = find_ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (only available in the build container)")
    for name in ("av", "decord"):
        try:
            __import__(name)
        except ImportError:
            sys.modules[name] = types.ModuleType(name)
    import mantis  # noqa: F401  (the alias shim; submodules below are added next to it)
    if "mantis.train" not in sys.modules:
        m = types.ModuleType("mantis.train")
        m.__path__ = []
        sys.modules["mantis.train"] = m
    if "mantis.models.conversation" not in sys.modules:
        _load("mantis.models.conversation", os.path.join(root, "mantis", "models", "conversation.py"))
    for sub in ("train_utils", "conversation"):
        if f"mantis.train.{sub}" not in sys.modules:
            _load(f"mantis.train.{sub}", os.path.join(root, "mantis", "train", f"{sub}.py"))
    mod = _load("_mantis_ref_train_data", os.path.join(root, "mantis", "train", "data.py"))
    _CACHE["train_data"] = mod
    return mod


def ref_llava_classes():
    cfg, modm = load_reference_mllava()

    class RefLlava(modm.LlavaForConditionalGeneration):
        _supports_sdpa = True

        def tie_weights(self, *a, **k):
            return None

    class RefMLlava(modm.MLlavaForConditionalGeneration):
        _supports_sdpa = True

        def tie_weights(self, *a, **k):
            return None

    return cfg.LlavaConfig, RefLlava, RefMLlava


def ref_idefics2_classes():
    modm = load_reference_idefics2()

    class RefIdefics2(modm.Idefics2ForConditionalGeneration):
        _supports_sdpa = True

        def tie_weights(self, *a, **k):
            return None

    return RefIdefics2
