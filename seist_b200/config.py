"""`Config` registry surface of the reference (config.py) for the SeisT path.

What the reference's training worker asks of it (training/train.py:199-202,269-275): regex-keyed model
configurations (`Config.models`, config.py:64-186) giving the loss class, the input / label groups and the
evaluated tasks; `get_model_config_`, `get_num_inchannels`, `get_loss`, `get_metrics`, `get_num_classes`
(config.py:327-432).  Only the five `seist_*` families are configured here — the comparison models of the
reference are outside the accelerated path.  Matching and error behaviour follow the reference:
exactly one regex key must match a registered model name (config.py:362-372).
"""
import re
from collections import defaultdict
from functools import partial
from typing import Any

from .models import BCELoss, CELoss, HuberLoss, get_model_list


def _entry(loss, labels, evals):
    return {"loss": loss, "inputs": [["z", "n", "e"]], "labels": labels, "eval": evals,
            "targets_transform_for_loss": None, "outputs_transform_for_loss": None,
            "outputs_transform_for_results": None}


class Config:
    _model_conf_keys = ("loss", "labels", "eval", "outputs_transform_for_loss", "outputs_transform_for_results")

    models = {
        # detection + P/S picking: per-sample BCE, detection channel weighted 0.5 (config.py:137-145)
        "seist_.*?_dpk.*": _entry(partial(BCELoss, weight=[[0.5], [1], [1]]), [["det", "ppk", "spk"]],
                                  ["det", "ppk", "spk"]),
        "seist_.*?_pmp": _entry(partial(CELoss, weight=[1, 1]), ["pmp"], ["pmp"]),     # config.py:147-155
        "seist_.*?_emg": _entry(HuberLoss, ["emg"], ["emg"]),                           # config.py:157-165
        "seist_.*?_baz": _entry(HuberLoss, ["baz"], ["baz"]),                           # config.py:167-175
        "seist_.*?_dis": _entry(HuberLoss, ["dis"], ["dis"]),                           # config.py:177-185
    }

    _avl_metrics = ("precision", "recall", "f1", "mean", "rmse", "mae", "mape", "r2")
    _avl_io_item_types = ("soft", "value", "onehot")
    _wave = {"type": "soft", "metrics": ["mean", "rmse", "mae"]}
    _pick = {"type": "soft", "metrics": ["precision", "recall", "f1", "mean", "rmse", "mae", "mape"]}
    _scalar = {"type": "value", "metrics": ["mean", "rmse", "mae", "r2"]}
    _avl_io_items = {
        "z": _wave, "n": _wave, "e": _wave,
        "det": {"type": "soft", "metrics": ["precision", "recall", "f1"]},
        "ppk": _pick, "spk": _pick,
        "emg": _scalar, "baz": _scalar, "dis": _scalar,
        "pmp": {"type": "onehot", "metrics": ["precision", "recall", "f1"], "num_classes": 2},
    }

    @classmethod
    def check_and_init(cls):
        cls._type_to_ioitems = defaultdict(list)
        for k, v in cls._avl_io_items.items():
            if v["type"] not in cls._avl_io_item_types:
                raise NotImplementedError(f"Unknown item type: {v['type']}, item: {k}")
            if set(v["metrics"]) - set(cls._avl_metrics):
                raise NotImplementedError(f"Unknown metrics:{set(v['metrics']) - set(cls._avl_metrics)} , item: {k}")
            cls._type_to_ioitems[v["type"]].append(k)
        unused = [key for key in cls.models if not any(re.findall(key, n) for n in get_model_list())]
        if unused:
            print(f"Useless configurations: {unused}")
        for name, conf in cls.models.items():
            missing = set(cls._model_conf_keys) - set(conf)
            if missing:
                raise Exception(f"Model:'{name}'  Missing keys:{missing}")
            for field_ in ("labels", "inputs"):
                flat = sum([g if isinstance(g, (tuple, list)) else [g] for g in conf[field_]], [])
                if set(flat) - set(cls._avl_io_items):
                    raise NotImplementedError(f"Model:'{name}'  Unknown {field_}:{set(flat) - set(cls._avl_io_items)}")
            if set(conf["eval"]) - set(cls._avl_io_items):
                raise NotImplementedError(f"Model:'{name}'  Unknown tasks:{set(conf['eval']) - set(cls._avl_io_items)}")

    @classmethod
    def get_io_items(cls, type: str = None) -> list:
        return list(cls._avl_io_items) if type is None else cls._type_to_ioitems[type]

    @classmethod
    def get_type(cls, name: str) -> str:
        return cls._avl_io_items[name]["type"]

    @classmethod
    def get_num_classes(cls, name: str) -> int:
        if name not in cls._avl_io_items:
            raise ValueError(f"Name {name} not exists.")
        if cls._avl_io_items[name]["type"] != "onehot":
            raise Exception(f"Type of item '{name}' is '{cls._avl_io_items[name]['type']}'.")
        return cls._avl_io_items[name]["num_classes"]

    @classmethod
    def get_model_config(cls, model_name: str) -> dict:
        registered = get_model_list()
        if model_name not in registered:
            raise NotImplementedError(f"Unknown model:'{model_name}', registered: {registered}")
        keys = [k for k in cls.models if re.findall(k, model_name)]
        if len(keys) < 1:
            raise Exception(f"Missing configuration of model {model_name}")
        if len(keys) > 1:
            raise Exception(f"Model {model_name} matches multiple configuration items: {keys}")
        return cls.models[keys[0]]

    @classmethod
    def get_model_config_(cls, model_name: str, *attrs) -> Any:
        conf = cls.get_model_config(model_name)
        out = []
        for a in attrs:
            if a not in conf:
                raise Exception(f"Unknown attribute:'{a}', supported: {list(conf)}")
            out.append(conf[a])
        return out[0] if len(out) == 1 else tuple(out)

    @classmethod
    def get_num_inchannels(cls, model_name: str) -> int:
        for inp in cls.get_model_config_(model_name, "inputs"):
            if isinstance(inp, (list, tuple)) and cls._avl_io_items[inp[0]]["type"] == "soft":
                return len(inp)
        raise Exception(f"Incorrect input channels. Model:{model_name}")

    @classmethod
    def get_metrics(cls, item_name: str) -> list:
        if item_name not in cls._avl_io_items:
            raise Exception(f"Unknown item:'{item_name}', supported: {list(cls._avl_io_items)}")
        return cls._avl_io_items[item_name]["metrics"]

    @classmethod
    def get_loss(cls, model_name: str):
        return cls.get_model_config(model_name)["loss"]()


Config.check_and_init()
