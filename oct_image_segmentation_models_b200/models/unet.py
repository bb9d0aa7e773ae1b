"""U-Net model container: same constructor, `get_config`, `get_preprocess_input_fn` and
`build_model` surface as the reference class (models/unet.py:61-153); `build_model` returns a
GPU-backed model object instead of a Keras graph."""
from typing import Callable, Union

from .base_model import BaseModel
from .keras_like import B200Model

UNET_MODEL_NAME = "unet"


class UNet(BaseModel):
    def __init__(
        self,
        *,
        input_channels: int,
        num_classes: int,
        image_height: int,
        image_width: int,
        start_neurons: int = 8,
        pool_layers: int = 4,
        conv_layers: int = 2,
        enc_kernel: Union[list, tuple] = (3, 3),
        dec_kernel: Union[list, tuple] = (2, 2),
    ) -> None:
        super().__init__(
            input_channels=input_channels,
            num_classes=num_classes,
            image_height=image_height,
            image_width=image_width,
        )
        for name, val in (("start_neurons", start_neurons), ("pool_layers", pool_layers),
                          ("conv_layers", conv_layers)):
            if not isinstance(val, int) or isinstance(val, bool):
                raise TypeError(f"{name} must be int")
        if not isinstance(enc_kernel, (list, tuple)) or not isinstance(dec_kernel, (list, tuple)):
            raise TypeError("enc_kernel / dec_kernel must be list or tuple")
        self.start_neurons = start_neurons
        self.pool_layers = pool_layers
        self.conv_layers = conv_layers
        self.enc_kernel = tuple(enc_kernel)
        self.dec_kernel = tuple(dec_kernel)

    def get_preprocess_input_fn(self) -> Callable:
        def preprocess_input_inner(x):
            return x / 255.0

        return preprocess_input_inner

    def get_config(self) -> dict:
        config = super().get_config()
        config.update(
            {
                "start_neurons": self.start_neurons,
                "pool_layers": self.pool_layers,
                "conv_layers": self.conv_layers,
                "enc_kernel": self.enc_kernel,
                "dec_kernel": self.dec_kernel,
            }
        )
        return config

    def spec_kwargs(self) -> dict:
        return dict(input_channels=self.input_channels, num_classes=self.num_classes,
                    start_neurons=self.start_neurons, pool_layers=self.pool_layers,
                    conv_layers=self.conv_layers, enc_kernel=self.enc_kernel, dec_kernel=self.dec_kernel)

    def build_model(self, precision=None, device: int = 0) -> B200Model:
        return B200Model(UNET_MODEL_NAME, self.spec_kwargs(), precision=precision, device=device)
