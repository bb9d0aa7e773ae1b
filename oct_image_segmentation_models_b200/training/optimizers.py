"""Optimizer constructors for `TrainingParams.opt_con` (reference training/training.py:190-193:
`optimizer = optimizer_con(**optimizer_params)`).  The update itself is the fused Keras-formulation Adam kernel
(csrc/train_kernels.cu adam_kernel: epsilon outside the bias correction, as tf.keras optimizer_v2); this class only
carries the hyper-parameters, with the `get_config()` the reference logs (training.py:124-131)."""


class Adam:
    _name = "Adam"

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, name="Adam", **kwargs):
        if amsgrad:
            raise NotImplementedError("amsgrad is not implemented by the fused Adam kernel")
        if "lr" in kwargs:                       # Keras 2.x alias
            learning_rate = kwargs.pop("lr")
        unknown = set(kwargs) - {"decay", "clipnorm", "clipvalue", "global_clipnorm"}
        if unknown:
            raise TypeError(f"Adam got unexpected arguments {sorted(unknown)}")
        if any(kwargs.get(k) for k in ("decay", "clipnorm", "clipvalue", "global_clipnorm")):
            raise NotImplementedError("learning-rate decay / gradient clipping are not implemented by the fused Adam kernel")
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.name = name

    def get_config(self):
        return {"name": self.name, "learning_rate": self.learning_rate, "decay": 0.0, "beta_1": self.beta_1,
                "beta_2": self.beta_2, "epsilon": self.epsilon, "amsgrad": False}


def adam_hyperparameters(optimizer) -> dict:
    """Accepts an instance of the class above, the string 'adam', or any duck-typed Keras Adam (get_config())."""
    if optimizer is None or (isinstance(optimizer, str) and optimizer.lower() == "adam"):
        optimizer = Adam()
    cfg = optimizer.get_config() if hasattr(optimizer, "get_config") else None
    name = (cfg or {}).get("name", type(optimizer).__name__)
    if cfg is None or str(name).lower() != "adam":
        raise NotImplementedError(f"optimizer {name!r}: the accelerated train step implements Adam only")
    if cfg.get("amsgrad"):
        raise NotImplementedError("amsgrad is not implemented by the fused Adam kernel")
    lr = cfg.get("learning_rate", 1e-3)
    if isinstance(lr, dict):
        raise NotImplementedError("learning-rate schedules are not implemented by the fused Adam kernel")
    return {"learning_rate": float(lr), "beta_1": float(cfg.get("beta_1", 0.9)), "beta_2": float(cfg.get("beta_2", 0.999)),
            "epsilon": float(cfg.get("epsilon", 1e-7))}
