"""Mirror of the reference's ``utils`` package layout for the hot path, so that
``from style_transfer_based_holographic_imaging_b200.utils.Forward_model import Holo_Generator`` reads like
the reference's ``from utils.Forward_model import Holo_Generator`` (test_field_retrieval_mnist.py:27)."""
