"""Host-side adapters between the data batch and the grounding tokenizer (UniFusion)."""
