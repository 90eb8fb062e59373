"""Host logic of B200Trainer: LR schedules == the ones HF Trainer builds for the reference's launch script
(mantis/train/scripts/train_mllava.sh:162-165: cosine, warmup_ratio 0.03), optimizer-state checkpoint round trip."""
import math

import pytest
import torch

from mantis_b200.train.engine import B200Trainer, lr_lambda


@pytest.mark.parametrize("kind", ["cosine", "linear", "constant"])
def test_lr_schedule_matches_transformers(kind):
    from transformers import get_scheduler
    total, ratio, base = 200, 0.03, 1e-5
    warm = int(math.ceil(total * ratio))
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=base)
    name = {"constant": "constant_with_warmup"}.get(kind, kind)
    sched = get_scheduler(name, opt, num_warmup_steps=warm, num_training_steps=total)
    for step in range(total):
        hf = opt.param_groups[0]["lr"]
        ours = base * lr_lambda(step, total, warm, kind)
        assert abs(hf - ours) <= 1e-12 * base + 1e-18, (step, hf, ours)
        opt.step(); sched.step()


def test_trainer_schedule_and_state_roundtrip():
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    tr = B200Trainer(model, lr=2e-5, lr_schedule="cosine", total_steps=100, warmup_ratio=0.03, freeze_vision=False,
                     fused_wgrad_accum=False, overlap_allreduce=False)
    assert tr.warmup_steps == 3 and tr.current_lr() == 0.0
    tr.step_count = 3
    assert tr.current_lr() == pytest.approx(2e-5)
    tr.step_count = 100
    assert tr.current_lr() == pytest.approx(0.0, abs=1e-12)
    # weights, gradients and moments live in flat buffers; every slice starts on a 1024-element boundary
    st = tr.state
    assert all(p.data_ptr() == st.P.data_ptr() + 4 * o and o % 1024 == 0 for p, o in zip(tr.params, st.offsets))
    assert st.groups.tolist() == [0, 1, 0, 1]                 # weight, bias, weight, bias: biases take no weight decay
    tr.step_count = 17
    st.M.normal_(); st.V.uniform_()
    state = tr.state_dict()
    tr2 = B200Trainer(torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4)), freeze_vision=False,
                      fused_wgrad_accum=False, overlap_allreduce=False)
    tr2.load_state_dict(state)
    assert tr2.step_count == 17 and tr2.lr == 2e-5 and tr2.lr_schedule == "cosine" and tr2.warmup_steps == 3
    for i in range(len(tr.params)):
        assert torch.equal(st.view(st.M, i), tr2.state.view(tr2.state.M, i))
        assert torch.equal(st.view(st.V, i), tr2.state.view(tr2.state.V, i))
        assert torch.equal(tr.params[i], tr2.params[i])       # the fp32 master weights travel with the optimizer state
    assert tr2.current_lr() == tr.current_lr()
    with pytest.raises(ValueError):
        B200Trainer(torch.nn.Linear(4, 4), freeze_vision=False, fused_wgrad_accum=False,
                    overlap_allreduce=False).load_state_dict(state)
