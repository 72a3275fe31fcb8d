"""Config plug-in mechanism of the reference (utils/utils.py:27-42): YAML `target:` strings resolved with
importlib and called with `**params`.  This is the boundary through which virtual_render/virtual_pose_render.py
builds the model, so the same `target:` paths resolve to the B200-native classes of this repo."""
import importlib


def count_params(model, verbose=False):
    n = sum(p.numel() for p in model.parameters())
    if verbose:
        print(f"{type(model).__name__} has {n * 1e-6:.2f} M params.")
    return n


def check_istarget(name, para_list):
    return any(p in name for p in para_list)


def get_obj_from_str(string, reload=False):
    module_name, attr = string.rsplit(".", 1)
    module = importlib.import_module(module_name)
    if reload:
        module = importlib.reload(module)
    return getattr(module, attr)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    params = config.get("params", dict()) or dict()
    return get_obj_from_str(config["target"])(**params)
