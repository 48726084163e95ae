# evoworld.inference.forward_evoworld (forward_evoworld.py:119-211) -> evoworld_b200.inference
from evoworld_b200.inference import prepare_batch_data, process_batch, save_frames  # noqa: F401
