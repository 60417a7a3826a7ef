"""FDR classifier — drop-in for ``BinaryClassifierLegacyNewBatching`` (alphadia/fdr/classifiers.py:145-470), the object
``perform_fdr`` (alphadia/fdr/fdr.py:25-192) calls ``fit`` and ``predict_proba`` on.  SURVEY 8f.2.

Same constructor arguments, ``to_state_dict`` / ``from_state_dict`` (the reference's own dictionaries load unchanged, so a
classifier trained by the reference can be used here and vice versa), ``predict`` and ``predict_proba``.

* ``predict_proba`` / ``predict`` run the network (BatchNorm1d in eval mode -> [Linear -> ReLU] x 4 -> Linear -> softmax,
  classifiers.py:473-532) in ONE CUDA kernel behind ``adb_classifier_predict_proba``: float32 like torch, within 1e-4 of the
  reference's output for the same weights (tests/golden/classifier_small.npz).  No CPU fallback.
* ``fit`` trains the same architecture with the same recipe (Adam, BCE on the two softmax outputs, fixed-size batches whose
  ORDER is shuffled per epoch, the last partial batches dropped, classifiers.py:311-427) with torch on the CUDA device: the
  optimiser loop of a 12 k-parameter network is plumbing, not a hot path, and its result is not bit-reproducible across
  devices in the reference either.
"""

from __future__ import annotations

import ctypes as C
import logging
import warnings
from copy import deepcopy

import numpy as np

from alphadia_b200 import _abi, _lib

logger = logging.getLogger()

BN_EPS = 1e-5  # torch.nn.BatchNorm1d default, which the reference uses


class _ClassifierDesc(C.Structure):
    _fields_ = [
        ("input_dim", C.c_int32), ("n_layers", C.c_int32), ("layer_dims", C.POINTER(C.c_int32)),
        ("bn_weight", _abi.c_f32p), ("bn_bias", _abi.c_f32p), ("bn_mean", _abi.c_f32p), ("bn_var", _abi.c_f32p),
        ("bn_eps", C.c_float), ("weights", C.POINTER(_abi.c_f32p)), ("biases", C.POINTER(_abi.c_f32p)),
    ]


def _to_numpy(v) -> np.ndarray:
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(v), dtype=np.float32)


def linear_layer_keys(n_linear: int) -> list[str]:
    """Module names of the Linear layers inside ``FeedForwardNN.fc_layers`` (BatchNorm1d is module 0; every hidden layer is
    Linear, ReLU, Dropout; classifiers.py:506-521)."""
    return [f"fc_layers.{1 + 3 * i}" for i in range(n_linear)]


def network_forward_device(network_state: dict, x: np.ndarray, device: int | None = None) -> np.ndarray:
    """softmax(FeedForwardNN(x)) in eval mode on the device for a ``network.state_dict()``-shaped dictionary."""
    _lib.require_device()
    lib = _lib.load()
    n_linear = sum(1 for k in network_state if k.endswith(".weight")) - 1  # minus the batch norm
    keys = linear_layer_keys(n_linear)
    weights = [_to_numpy(network_state[k + ".weight"]) for k in keys]
    biases = [_to_numpy(network_state[k + ".bias"]) for k in keys]
    bn = {k: _to_numpy(network_state["fc_layers.0." + k]) for k in ("weight", "bias", "running_mean", "running_var")}
    x = np.ascontiguousarray(x, dtype=np.float32)  # torch.Tensor(x) is float32 (classifiers.py:469)
    input_dim = int(weights[0].shape[1])
    if x.ndim != 2 or x.shape[1] != input_dim:
        raise ValueError(f"x must be [n_samples, {input_dim}]")
    dims = (C.c_int32 * n_linear)(*[int(w.shape[0]) for w in weights])
    w_ptrs = (_abi.c_f32p * n_linear)(*[_abi.ptr(w) for w in weights])
    b_ptrs = (_abi.c_f32p * n_linear)(*[_abi.ptr(b) for b in biases])
    d = _ClassifierDesc(input_dim, n_linear, dims, _abi.ptr(bn["weight"]), _abi.ptr(bn["bias"]), _abi.ptr(bn["running_mean"]),
                        _abi.ptr(bn["running_var"]), BN_EPS, w_ptrs, b_ptrs)
    out = np.empty((x.shape[0], int(weights[-1].shape[0])), dtype=np.float32)
    dev = _lib.current_device() if device is None else device
    _lib.check(lib.adb_classifier_predict_proba(C.c_int(dev), C.byref(d), C.c_int64(x.shape[0]), _abi.ptr(x), _abi.ptr(out)),
               "adb_classifier_predict_proba")
    return out


class BinaryClassifierLegacyNewBatching:
    """Binary target / decoy classifier, feed-forward network (classifiers.py:145-470)."""

    def __init__(self, input_dim: int = 10, output_dim: int = 2, test_size: float = 0.2, batch_size: int = 1000, epochs: int = 10,
                 learning_rate: float = 0.0002, weight_decay: float = 0.00001, layers: list[int] | None = None,
                 dropout: float = 0.001, metric_interval: int = 1000, *, experimental_hyperparameter_tuning: bool = False,
                 random_state: int | None = None, **kwargs):
        self.test_size = test_size
        self.batch_size = batch_size
        self.epochs = epochs
        self.learning_rate = learning_rate
        self.weight_decay = weight_decay
        self.layers = [100, 50, 20, 5] if layers is None else layers
        self.dropout = dropout
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.metric_interval = metric_interval
        self.experimental_hyperparameter_tuning = experimental_hyperparameter_tuning
        self.network_state = None  # network.state_dict() as numpy arrays
        self._fitted = False
        self.metrics = {"epoch": [], "batch_count": [], "train_loss": [], "train_accuracy": [], "test_loss": [], "test_accuracy": []}
        self._np_rng = np.random.default_rng(seed=random_state)
        self._torch_seed = int(self._np_rng.integers(0, 1_000_000)) if random_state is not None else None
        if kwargs:
            warnings.warn(f"Unknown arguments: {kwargs}")

    @property
    def fitted(self) -> bool:
        return self._fitted

    # ---- persistence: the reference's dictionaries (classifiers.py:258-309) -----------------------------------------------
    def to_state_dict(self) -> dict:
        state = {k: getattr(self, k) for k in ("input_dim", "output_dim", "test_size", "batch_size", "epochs", "learning_rate",
                                              "weight_decay", "layers", "dropout", "metric_interval", "metrics")}
        state["_fitted"] = self._fitted
        if self._fitted:
            state["network_state_dict"] = {k: v.copy() for k, v in self.network_state.items()}
        return state

    def from_state_dict(self, state_dict: dict, *, load_hyperparameters: bool = False) -> None:
        state = deepcopy({k: v for k, v in state_dict.items() if k != "network_state_dict"})
        if "network_state_dict" in state_dict:
            net = state_dict["network_state_dict"]
            self.network_state = {k: (_to_numpy(v) if not k.endswith("num_batches_tracked") else np.asarray(
                v.detach().cpu().numpy() if hasattr(v, "detach") else v)) for k, v in net.items()}
            self.input_dim = int(state.pop("input_dim"))
            self.output_dim = int(state.pop("output_dim"))
            self.layers = list(state.pop("layers"))
            self.dropout = state.pop("dropout")
            self._fitted = True
        if load_hyperparameters:
            self.__dict__.update(state)

    # ---- inference ------------------------------------------------------------------------------------------------------
    def _check_input(self, x: np.ndarray) -> None:
        if not self.fitted:
            raise ValueError("Classifier has not been fitted yet.")
        assert x.ndim == 2, "Input data must have batch and feature dimension. (n_samples, n_features)"  # noqa: PLR2004
        assert x.shape[1] == self.input_dim, "Input data must have the same number of features as the fitted classifier."

    def predict_proba(self, x: np.ndarray) -> np.ndarray:
        """Class probabilities ``[n_samples, 2]`` (classifiers.py:441-470)."""
        self._check_input(x)
        return network_forward_device(self.network_state, x)

    def predict(self, x: np.ndarray) -> np.ndarray:
        """Predicted class (classifiers.py:412-439)."""
        return np.argmax(self.predict_proba(x), axis=1)

    # ---- training (classifiers.py:311-427) ---------------------------------------------------------------------------------
    def fit(self, x: np.ndarray, y: np.ndarray) -> None:
        import torch
        from sklearn.model_selection import train_test_split
        from torch import nn

        _lib.require_device()
        dev = torch.device("cuda", _lib.current_device())
        if self.experimental_hyperparameter_tuning:
            # classifiers.py:104-143 _get_scaled_training_params: batch size linear in the sample count between 128 and 4096 (at
            # one million samples), learning rate scaled with the square root of the batch size
            base_lr, max_batch, min_batch = 0.001, 4096, 128
            if len(x) >= 1_000_000:
                self.batch_size, self.learning_rate = max_batch, base_lr
            else:
                self.batch_size = int(np.clip((len(x) / 1_000_000) * max_batch, min_batch, max_batch))
                self.learning_rate = float(base_lr * np.sqrt(self.batch_size / max_batch))
        if self.network_state is not None and self.input_dim != x.shape[1]:
            warnings.warn("Input dimension of network has changed. Network has been reinitialized.")
            self.network_state = None
        self.input_dim = x.shape[1]
        if self._torch_seed is not None:
            torch.manual_seed(self._torch_seed)
        widths = [self.input_dim, *self.layers]
        mods = [nn.BatchNorm1d(self.input_dim)]
        for a, b in zip(widths[:-1], widths[1:]):
            mods += [nn.Linear(a, b), nn.ReLU(), nn.Dropout(self.dropout)]
        mods += [nn.Linear(widths[-1], self.output_dim), nn.Softmax(dim=1)]
        net = nn.Sequential(*mods)
        if self.network_state is not None:  # continue from the current weights, as the reference keeps its network
            net.load_state_dict({k[len("fc_layers."):]: torch.as_tensor(v) for k, v in self.network_state.items()})
        net = net.to(dev)
        if y.ndim == 1:
            y = np.stack([1 - y, y], axis=1)
        split_seed = int(self._np_rng.integers(0, 1_000_000))
        x_train, x_test, y_train, y_test = train_test_split(x, y, test_size=self.test_size, random_state=split_seed)
        x_train, y_train = torch.tensor(x_train, dtype=torch.float32, device=dev), torch.tensor(y_train, dtype=torch.float32, device=dev)
        x_test, y_test = torch.tensor(x_test, dtype=torch.float32, device=dev), torch.tensor(y_test, dtype=torch.float32, device=dev)
        opt = torch.optim.Adam(net.parameters(), lr=self.learning_rate, weight_decay=self.weight_decay)
        loss_fn = nn.BCELoss()
        net.train()
        n_batches = (x_train.shape[0] // self.batch_size) - 1
        starts = np.arange(max(n_batches, 0)) * self.batch_size
        count = 0
        for epoch in range(self.epochs):
            for s in starts[self._np_rng.permutation(len(starts))]:
                xb, yb = x_train[s:s + self.batch_size], y_train[s:s + self.batch_size]
                loss = loss_fn(net(xb), yb)
                opt.zero_grad()
                loss.backward()
                opt.step()
                if count % self.metric_interval == 0:
                    net.eval()
                    with torch.no_grad():
                        p_test, p_train = net(x_test), net(xb)
                        self.metrics["epoch"].append(epoch)
                        self.metrics["batch_count"].append(count)
                        self.metrics["train_loss"].append(float(loss.item()))
                        self.metrics["test_loss"].append(float(loss_fn(p_test, y_test).item()))
                        self.metrics["train_accuracy"].append(float((p_train.argmax(1) == yb[:, 1]).float().mean().item()))
                        self.metrics["test_accuracy"].append(float((p_test.argmax(1) == y_test[:, 1]).float().mean().item()))
                    net.train()
                count += 1
        self.network_state = {"fc_layers." + k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
        self._fitted = True
