"""Empty stand-in: clap_module/utils.py:6 imports h5py at module top; nothing on the audio path uses it."""
