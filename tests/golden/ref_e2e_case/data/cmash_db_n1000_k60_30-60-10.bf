stand-in for the hydra Bloom prefilter
