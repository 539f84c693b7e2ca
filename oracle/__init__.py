"""ORACLE — test infrastructure, not product code.  PARITY UNPINNED (the reference ships no tests,
fixtures or golden vectors; JAX/Flax are not installable here — SURVEY.md §8c).

CPU fp32 restatement of the reference's captioning hot path:
  reference_model.py     forward / loss / grads / AdamW   (modeling_clip_vision_mbart.py, main.py)
  reference_generate.py  greedy + beam search + cached decode (generation_clip_vision_utils.py)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package (`multilingual-image-captioning_b200`) must never import it.
"""
