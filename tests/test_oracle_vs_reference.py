"""Live check of the oracle against the unmodified reference — only where /root/reference exists
(the build container). Skipped on the GPU box."""
import numpy as np
import pytest
import torch

from mmtg_b200 import synth
from mmtg_b200.configs import data_config
from oracle import mmtg_oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")


def test_oracle_matches_reference_live():
    torch.set_num_threads(8)
    table = synth.make_token_table()
    sd = synth.make_state_dict(3)
    model, crit, _gen, _dc, _cfgs = ref_import.load_reference(table, sd)
    assert list(model.state_dict().keys()) == synth.state_dict_keys()
    batch = synth.batch_to_torch(synth.make_batch(3, seed=5))
    with torch.no_grad():
        hf, kl, logits = model(batch)
        ohf, okl, ologits = O.mmtg_forward(sd, torch.from_numpy(table), batch, data_config(), True)
        assert (logits - ologits).abs().max().item() < 2e-5
        assert abs(hf.item() - ohf.item()) < 1e-5 and abs(kl.item() - okl.item()) < 1e-5
        for stage in (1, 2, 3):
            a = crit(logits, batch["targets"], batch["rating"], stage).item()
            b = O.my_loss(ologits, batch["targets"], batch["rating"], stage).item()
            assert abs(a - b) < 1e-5 * max(1.0, abs(a))
