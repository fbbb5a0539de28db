class Template:
    def __init__(self, text=None, **kwargs):
        self.source = text

    def render(self, **kwargs):
        raise NotImplementedError("mako is not available offline (shim)")
