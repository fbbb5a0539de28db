class DeepDiff(dict):
    def __init__(self, old, new, **kwargs):
        super().__init__()
        if old != new:
            self["values_changed"] = {"root": {"old_value": old, "new_value": new}}
