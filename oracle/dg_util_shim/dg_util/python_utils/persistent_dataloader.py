"""Import-level stand-in for dg_util.python_utils.persistent_dataloader (TEST INFRASTRUCTURE ONLY)."""
import torch.utils.data


class PersistentDataLoader(torch.utils.data.DataLoader):
    def __init__(self, *args, never_ending=False, **kwargs):
        kwargs.pop("device", None)
        super().__init__(*args, **kwargs)
