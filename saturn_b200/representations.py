"""Task / HParams / Strategy / Techniques — the objects that carry the profiled runtime table T
into the solver and the selected plan out of it.

Mirror of the reference's public data model (saturn/core/representations/Task.py:23-179,
Strategy.py:25-76): same class names, constructor arguments, attributes and error behaviour, so
user scripts and the reference's own solver/executor code keep working on these objects.  The
implementation is independent; torch is imported lazily and Ray is not needed at all.

What the solver path touches:
  Task.strategies        dict gpu_count -> Strategy, insertion-ordered   (Task.py:118, milp.py:77-81)
  Strategy.runtime       estimated seconds for the remaining batches     (Strategy.py:73)
  Task.select_strategy   called by convert_into_comprehensible           (Task.py:171, milp.py:481-486)
  Task.total_batches     consumed / decremented by forecast              (Task.py:127-128, executor.py:166-172)
"""
from __future__ import annotations

import enum
import os
import secrets
import string
from typing import Callable, Dict, List, Optional

_ALPHABET = string.ascii_uppercase + string.digits


def _random_name(n: int = 16) -> str:
    return "".join(secrets.choice(_ALPHABET) for _ in range(n))


class Techniques(enum.Enum):
    """Families of parallelism a Strategy's executor may belong to (Strategy.py:25-34)."""
    SPILLED = 1
    PIPELINE = 2
    FSDP = 3
    MEGATRON = 4


class Strategy:
    """One (executor, GPU count) choice for a task together with its profiled runtime."""

    def __init__(self, executor, gpu_apportionment: int, parameters: Optional[dict] = None, runtime=None) -> None:
        if isinstance(gpu_apportionment, bool) or not isinstance(gpu_apportionment, int) or gpu_apportionment <= 0:
            raise ValueError("GPU allocation must be an integer > 0.")
        self.executor = executor
        self.gpu_apportionment = gpu_apportionment
        self.parameters = parameters
        self.runtime = runtime

    def __repr__(self) -> str:
        return "Strategy({} ({}), {}G, {}s)".format(self.executor, self.parameters, self.gpu_apportionment,
                                                    self.runtime)

    __str__ = __repr__


class HParams:
    """Training hyper-parameters of a task; exactly one of `epochs` / `batch_count` must be given."""

    def __init__(self, lr: float, epochs: Optional[int] = None, batch_count: Optional[int] = None,
                 optimizer_cls=None, **kwargs) -> None:
        if (epochs is None) == (batch_count is None):
            raise ValueError("Exactly one of epochs and batch_count must be set (got epochs={!r}, "
                             "batch_count={!r}).".format(epochs, batch_count))
        self.lr = lr
        self.epochs = epochs
        self.batch_count = batch_count
        self.optimizer_cls = optimizer_cls
        self.kwargs = kwargs

    def as_dict(self) -> dict:
        return {"lr": self.lr, "epochs": self.epochs, "batch_count": self.batch_count}

    def __str__(self) -> str:
        tail = "Epochs: {}".format(self.epochs) if self.epochs is not None else "Batch Count: {}".format(
            self.batch_count)
        return "----Task Hyperparameters----\n\t\tLearning Rate: {}\n\t\t{}".format(self.lr, tail)


class Task:
    """A model-training job: model / dataloader factories, loss, hyper-parameters, and — after
    profiling — the table row `strategies` the solver consumes."""

    def __init__(self, get_model: Callable, get_dataloader: Callable, loss_function: Callable, hparams: HParams,
                 gpu_range: Optional[List[int]] = None, name: Optional[str] = None, hints: Optional[dict] = None,
                 save_dir: str = "./saved_models"):
        if hints is not None and hints.get("is_transformer", False):
            cls = hints.get("transformer_cls", None)
            if cls is None or not isinstance(cls, set):
                raise ValueError("A task flagged is_transformer must pass its attention-block classes as a set "
                                 "under hints['transformer_cls'].")
        self.hints = hints
        self.internal_get_model = get_model
        self.internal_dl = get_dataloader
        self.hparams = hparams
        self.loss_function = loss_function
        self.gpu_range = gpu_range
        self.name = name if name is not None else _random_name()
        self.saved_dataloader = None
        self.save_dir = save_dir
        os.makedirs(save_dir, exist_ok=True)
        self.strategies: Dict[int, Strategy] = {}
        self.selected_strategy: Optional[Strategy] = None
        self.epoch_length = len(self.internal_dl())
        if self.hparams.epochs:
            self.total_batches = self.epoch_length * self.hparams.epochs
        else:
            self.total_batches = self.hparams.batch_count
        self.current_batch = 0

    # -- data position ---------------------------------------------------------------------
    def get_iterator(self, modified_dl=None):
        it = iter(modified_dl if modified_dl is not None else self.internal_dl())
        for _ in range(self.current_batch):
            next(it)
        return it

    def get_fresh_iterator(self):
        return iter(self.internal_dl())

    def reconfigure(self, batch_count: int) -> None:
        self.current_batch = (self.current_batch + batch_count) % self.epoch_length

    def change_name(self, name: Optional[str] = None) -> None:
        # the reference only ever assigns a fresh random name here (Task.py:145-148)
        if name is None:
            self.name = _random_name()

    # -- checkpoints -----------------------------------------------------------------------
    def _ckpt_path(self) -> str:
        return "{}/{}.pt".format(self.save_dir, self.name)

    def has_ckpt(self) -> bool:
        return os.path.isfile(self._ckpt_path())

    def save(self, model) -> None:
        import torch
        print("Saving model {}/{}".format(self.save_dir, self.name))
        torch.save(model, self._ckpt_path())
        print("Saved model {}/{}".format(self.save_dir, self.name))

    def get_model(self, fresh: bool = False):
        if self.has_ckpt() and not fresh:
            import torch
            return torch.load(self._ckpt_path())
        if self.hparams.kwargs:
            return self.internal_get_model(self.hparams.kwargs)
        return self.internal_get_model()

    # -- plan ------------------------------------------------------------------------------
    def select_strategy(self, strat: Strategy) -> None:
        self.selected_strategy = strat

    def __str__(self) -> str:
        return ("----Task {}----\n\t{}\n\tCandidate Strategies: {}\n\tSelected Strategy: {}\n".format(
            self.name, self.hparams, self.strategies, self.selected_strategy))
