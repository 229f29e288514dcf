from gpt_b200.qcd import fermion, gauge
