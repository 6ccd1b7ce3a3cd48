"""CPU oracle for the halo2 MSM+FFT hot path: TEST INFRASTRUCTURE ONLY (see pasta.py header)."""
