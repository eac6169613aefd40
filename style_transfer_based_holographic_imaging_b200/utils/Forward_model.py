from ..Forward_model import Holo_Generator, Back_prop, center_crop  # noqa: F401
