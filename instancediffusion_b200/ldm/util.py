"""Config-driven construction, the reference's only plug-in mechanism (ldm/util.py:71-84):
YAML `target:` dotted paths resolved with importlib.  With instancediffusion_b200.dropin.install()
the reference's own paths (`ldm.modules...`) resolve to the classes of this package."""
import importlib


def get_obj_from_str(string, reload=False):
    module_name, cls_name = string.rsplit(".", 1)
    module = importlib.import_module(module_name)
    if reload:
        module = importlib.reload(module)
    return getattr(module, cls_name)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))


def default(val, d):
    if val is not None:
        return val
    return d() if callable(d) else d
