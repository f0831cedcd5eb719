"""Hyper-parameter containers consumed by the drop-in modules.

Field names and defaults follow the reference (trphysx/config/configuration_phys.py:52-74,
configuration_lorenz.py:23-29, configuration_cylinder.py:23-27, examples/rossler/rossler_module/
configuration_rossler.py:24-30).  The modules only read attributes, so a reference `PhysConfig`
instance can be passed wherever one of these is expected.
"""
import copy
import json
import os
from collections import OrderedDict

CONFIG_NAME = "config_trphysx.json"
_DEFAULTS = OrderedDict(
    activation_function="gelu_new", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=1e-5,
    initializer_range=0.01, output_hidden_states=False, output_attentions=False, use_cache=True,
    max_length=8, min_length=0, k_size=1)


class PhysConfig:
    model_type = ""

    def __init__(self, **kwargs):
        for req in ("n_ctx", "n_embd", "n_layer", "n_head", "state_dims"):
            setattr(self, req, kwargs.pop(req))          # KeyError when missing, like the reference
        for key, default in _DEFAULTS.items():
            setattr(self, key, kwargs.pop(key, default))
        for key, value in kwargs.items():               # free-form extras are kept as attributes
            setattr(self, key, value)

    # properties shared by every concrete config of the reference
    @property
    def hidden_size(self):
        return self.n_embd

    @property
    def num_attention_heads(self):
        return self.n_head

    @property
    def num_hidden_layers(self):
        return self.n_layer

    def to_dict(self):
        out = copy.deepcopy(self.__dict__)
        out["model_type"] = type(self).model_type
        return out

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def to_json_file(self, path):
        with open(path, "w", encoding="utf-8") as f:
            f.write(self.to_json_string())

    def save_pretrained(self, save_directory):
        if os.path.isfile(save_directory):
            raise AssertionError(f"Provided path ({save_directory}) should be a directory, not a file")
        os.makedirs(save_directory, exist_ok=True)
        self.to_json_file(os.path.join(save_directory, CONFIG_NAME))

    @classmethod
    def from_dict(cls, config_dict, **kwargs):
        cfg = cls(**{k: v for k, v in config_dict.items() if k != "model_type"})
        for k, v in kwargs.items():
            if hasattr(cfg, k):
                setattr(cfg, k, v)
        return cfg

    def update(self, config_dict):
        for k, v in config_dict.items():
            setattr(self, k, v)


def _make(name, model_type, **defaults):
    def __init__(self, **kwargs):
        merged = dict(defaults)
        merged.update(kwargs)
        PhysConfig.__init__(self, **merged)

    return type(name, (PhysConfig,), {"model_type": model_type, "__init__": __init__})


LorenzConfig = _make("LorenzConfig", "lorenz", n_ctx=64, n_embd=32, n_layer=4, n_head=4, state_dims=[3],
                     activation_function="gelu_new", initializer_range=0.05)
RosslerConfig = _make("RosslerConfig", "rossler", n_ctx=32, n_embd=32, n_layer=4, n_head=4, state_dims=[3],
                      activation_function="gelu_new", initializer_range=0.02)
CylinderConfig = _make("CylinderConfig", "cylinder", n_ctx=16, n_embd=128, n_layer=6, n_head=4,
                       state_dims=[3, 64, 128], activation_function="gelu_new")

# configuration_grayscott.py:15-38 (its model_type really is "cylinder" in the reference)
GrayScottConfig = _make("GrayScottConfig", "cylinder", n_ctx=128, n_embd=512, n_layer=2, n_head=32,
                        state_dims=[2, 32, 32, 32], activation_function="gelu_new")

CONFIG_MAPPING = OrderedDict([("lorenz", LorenzConfig), ("rossler", RosslerConfig), ("cylinder", CylinderConfig),
                              ("grayscott", GrayScottConfig)])


class AutoPhysConfig:
    """name -> config factory (reference: trphysx/config/configuration_auto.py:29-72)."""

    def __init__(self):
        raise EnvironmentError("AutoPhysConfig should not be initiated directly. Use its class methods.")

    @classmethod
    def load_config(cls, model_name_or_path, **kwargs):
        if model_name_or_path in CONFIG_MAPPING:
            return CONFIG_MAPPING[model_name_or_path](**kwargs)
        if os.path.isdir(model_name_or_path):
            model_name_or_path = os.path.join(model_name_or_path, CONFIG_NAME)
        if os.path.isfile(model_name_or_path):
            with open(model_name_or_path, "r", encoding="utf-8") as f:
                d = json.load(f)
            mt = d.pop("model_type", "")
            base = CONFIG_MAPPING.get(mt, PhysConfig)
            return base.from_dict(d, **kwargs)
        raise ValueError(f"Provided model name: {model_name_or_path} not present in built in configs or as a file")
