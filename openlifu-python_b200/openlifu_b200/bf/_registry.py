"""``{"class": name, **params}`` (de)serialisation shared by the bf method families
(reference: bf/delay_methods/delaymethod.py:21-32 and siblings)."""
from __future__ import annotations


class ClassKeyed:
    _family: dict

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        for base in cls.__mro__[1:]:
            fam = base.__dict__.get("_family")
            if fam is not None:
                fam[cls.__name__] = cls
                break

    def to_dict(self):
        d = {k: v for k, v in self.__dict__.items()}
        d["class"] = type(self).__name__
        return d

    @staticmethod
    def _from_dict(family_root, d):
        d = dict(d)
        name = d.pop("class")
        return family_root._family[name](**d)


def table(records):
    import pandas as pd
    return pd.DataFrame.from_records(records)
