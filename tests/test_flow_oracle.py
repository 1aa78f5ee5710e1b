"""CPU: flow-composition restatement vs the reference-generated fixtures."""
import torch

from conftest import assert_close, golden_cases, load_golden
from oracle import flow as FL


def test_flow_compose_oracle_matches_reference_fixtures():
    blob = load_golden("flow_compose")
    for name in [n for n in golden_cases(blob) if n.startswith("case")]:
        s = int(blob[name + ".s"])
        low, mask, base = [blob[name + k].clone().requires_grad_(True) for k in (".low", ".mask", ".base")]
        ident = FL.identity_flow_ref(s * low.shape[1], s * low.shape[2])
        delta, flow = FL.flow_compose_ref(low, mask, ident, base, None, s)
        assert_close(delta, blob[name + ".delta"], rtol=1e-6, what=name + " delta")
        assert_close(flow, blob[name + ".flow"], rtol=1e-6, what=name + " flow")
        grads = torch.autograd.grad((delta * blob[name + ".gd"]).sum() + (flow * blob[name + ".gf"]).sum(), [low, mask, base])
        for g, k in zip(grads, (".g_low", ".g_mask", ".g_base")):
            assert_close(g, blob[name + k], rtol=1e-5, what=name + k)


def test_similarity_matrix_oracle():
    blob = load_golden("flow_compose")
    m = FL.similarity_matrix_ref(blob["sim.params"])
    assert_close(m, blob["sim.matrix"], rtol=1e-6)
    assert_close(FL.compose_similarity_ref(blob["sim.base"], m), blob["sim.composed"], rtol=1e-6)
