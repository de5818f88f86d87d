"""EvalLanes (lanes.py): two batches in flight on two CUDA streams give bit-identical results to the plain loop."""
import pytest
import torch

from sound_event_detection_transformer_b200 import spec, synth
from sound_event_detection_transformer_b200.lanes import EvalLanes
from sound_event_detection_transformer_b200.sedt import build_model

pytestmark = pytest.mark.gpu


def _model(args, sd, graph):
    m, _, _ = build_model(args)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.use_cuda_graph = graph
    return m


@pytest.mark.parametrize("graph", [False, True])
def test_two_lanes_match_single_stream(graph):
    args = spec.config_args("c2")
    args.precision = "bf16"
    sd = synth.synth_state_dict(args, 12)
    batches = [synth.synth_clips(8, 496, 64, seed=40 + i).cuda() for i in range(6)]
    ref_model = _model(args, sd, graph)
    with torch.no_grad():
        want = [{k: v.clone() for k, v in ref_model(x).items() if torch.is_tensor(v)} for x in batches]
    lanes = EvalLanes([_model(args, sd, graph), _model(args, sd, graph)])
    got = []
    with torch.no_grad():
        for rep in range(2):                      # second round replays the captured graphs
            got.clear()
            lanes.fork()
            for i, x in enumerate(batches):
                with lanes.stream(i):
                    o = lanes.model(i)(x)
                    got.append({k: v.clone() for k, v in o.items() if torch.is_tensor(v)})
            lanes.join()
            torch.cuda.synchronize()
            for w, g in zip(want, got):
                for k in ("pred_logits", "pred_boxes", "at"):
                    assert torch.equal(w[k], g[k]), (rep, k)


def test_lanes_reject_shared_or_training_models():
    args = spec.config_args("c2")
    sd = synth.synth_state_dict(args, 12)
    m = _model(args, sd, False)
    with pytest.raises(ValueError):
        EvalLanes([m, m])
    with pytest.raises(ValueError):
        EvalLanes([])
    m.train()
    with pytest.raises(ValueError):
        EvalLanes([m])
