"""Stub of gensim.models for importing the reference's walk path (Word2Vec unused)."""


class Word2Vec:  # pragma: no cover - never instantiated by the golden generator
    def __init__(self, *a, **k):
        raise RuntimeError("gensim stub: Word2Vec is not available in this image")
