"""Config plumbing mirrored from the reference so its presets / YAML overrides apply unchanged.
ref: src/configs/configs.py:55-63 (InstantiateConfig), :214-242 (update_config),
     src/engine/callbacks.py (TrainingCallback*)
"""
from dataclasses import dataclass
from enum import Enum, auto
from inspect import signature
from typing import Any, Callable, Dict, List, Optional, Tuple, Type


@dataclass
class InstantiateConfig:
    """`XConfig(_target=X).setup(**kw)` instantiates X(config, **kw) (configs.py:55-63)."""

    _target: Type

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


def update_config(cfg, update: dict):
    """Recursive scalar/dict override from a YAML dict; same rules as Config.update_config (configs.py:214-242)."""

    def set_attribute(target, upd):
        for key in upd.keys():
            if isinstance(target, dict):
                cur = target.get(key)
                if hasattr(cur, "__dict__") or isinstance(cur, dict):
                    set_attribute(cur, upd[key])
                else:
                    target[key] = upd[key]
            else:
                cur = getattr(target, key)
                if hasattr(cur, "__dict__") or isinstance(cur, dict):
                    set_attribute(cur, upd[key])
                else:
                    setattr(target, key, upd[key])

    for key in update:
        if key in cfg.__dict__:
            target = getattr(cfg, key)
            if hasattr(target, "__dict__") or isinstance(target, dict):
                set_attribute(target, update[key])
            else:
                setattr(cfg, key, update[key])
    return cfg


class TrainingCallbackLocation(Enum):
    BEFORE_TRAIN_ITERATION = auto()
    AFTER_TRAIN_ITERATION = auto()


@dataclass
class TrainingCallbackAttributes:
    model: Any
    trainer: Any


class TrainingCallback:
    """ref: engine/callbacks.py:47-106"""

    def __init__(self, where_to_run: List[TrainingCallbackLocation], func: Callable,
                 update_every_num_iters: Optional[int] = None, iters: Optional[Tuple[int, ...]] = None,
                 args: Optional[List] = None, kwargs: Optional[Dict] = None):
        assert "step" in signature(func).parameters.keys(), "'step: int' must be an argument of the callback"
        self.where_to_run = where_to_run
        self.update_every_num_iters = update_every_num_iters
        self.iters = iters
        self.func = func
        self.args = args if args is not None else []
        self.kwargs = kwargs if kwargs is not None else {}

    def run_callback(self, step: int):
        if self.update_every_num_iters is not None:
            if step % self.update_every_num_iters == 0:
                self.func(*self.args, **self.kwargs, step=step)
        elif self.iters is not None and step in self.iters:
            self.func(*self.args, **self.kwargs, step=step)

    def run_callback_at_location(self, step: int, location: TrainingCallbackLocation):
        if location in self.where_to_run:
            self.run_callback(step=step)
