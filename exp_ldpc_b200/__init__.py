"""exp_ldpc_b200 — B200-native batched qLDPC decoding engine (drop-in for the qldpc decode path)."""
