"""Helper process of tests/test_model_parity_gpu.py::test_decode_engine_bitexact_with_and_without_pdl: runs a seeded bf16 LLaMA
through prefill + native decode steps and writes the logits to argv[1].  MB200_PDL is read once per process by the library, so
the two launch modes need two processes."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])


def main(out_path):
    from transformers import LlamaConfig
    from mantis_b200.models.decode_engine import native_decode_logits
    from mantis_b200.models.kv_cache import B200KVCache
    from mantis_b200.models.llama import B200CausalLM
    dev = torch.device("cuda:0")
    cfg = LlamaConfig(hidden_size=1024, intermediate_size=2816, num_hidden_layers=4, num_attention_heads=8,
                      num_key_value_heads=2, vocab_size=2000, rms_norm_eps=1e-5, rope_theta=500000.0)
    torch.manual_seed(1234)
    model = B200CausalLM(cfg).to(dev).to(torch.bfloat16).eval()
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(2.0)
    g = torch.Generator().manual_seed(99)
    ids = torch.randint(0, 2000, (2, 300), generator=g).to(dev)
    logits = []
    with torch.no_grad():
        cache = B200KVCache()
        out = model(input_ids=ids, past_key_values=cache, use_cache=True)
        nxt = out.logits[:, -1].argmax(-1)
        for _ in range(12):
            lg = native_decode_logits(model.model, model.lm_head, cache, nxt[:, None], torch.bfloat16, None, None)
            assert lg is not None, "decode engine not eligible"
            logits.append(lg.float().cpu().clone())
            nxt = lg.argmax(-1)
    torch.save(torch.stack(logits), out_path)


if __name__ == "__main__":
    main(sys.argv[1])
