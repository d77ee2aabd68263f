"""Stand-in package for the un-vendored torchlibrosa dependency (see stft.py). Test infrastructure only."""
