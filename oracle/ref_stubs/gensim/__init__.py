"""Stub: gensim is absent in this image; the walk path never touches it."""
