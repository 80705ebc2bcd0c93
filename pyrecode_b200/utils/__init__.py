"""Offline converters on reconstructed frames (drop-in for pyrecode/utils/converters.py)."""
