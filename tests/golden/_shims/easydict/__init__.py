class EasyDict(dict):
    """Attribute-access dict (stand-in for easydict.EasyDict, used by the reference's config.py:6)."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
