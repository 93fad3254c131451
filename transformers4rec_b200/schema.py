"""Minimal schema shim: just enough of ``merlin_standard_lib.Schema`` /
``merlin.schema.Tags`` for ``TabularSequenceFeatures.from_schema`` to keep its
signature (reference: merlin_standard_lib/schema/schema.py:215-550; cardinality =
``int_domain.max + 1`` at :541-550).  Host-side bookkeeping only."""
from __future__ import annotations

import enum
import json
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Union


class Tags(str, enum.Enum):
    CATEGORICAL = "categorical"
    CONTINUOUS = "continuous"
    LIST = "list"
    ITEM_ID = "item_id"
    ITEM = "item"
    USER_ID = "user_id"
    USER = "user"
    SESSION_ID = "session_id"
    EMBEDDING = "embedding"
    TIME = "time"
    TARGET = "target"
    BINARY_CLASSIFICATION = "binary_classification"
    REGRESSION = "regression"


TagsType = Union[Sequence[Union[str, Tags]], str, Tags]


def _norm_tags(tags) -> List[str]:
    if tags is None:
        return []
    if isinstance(tags, (str, Tags)):
        tags = [tags]
    return [t.value if isinstance(t, Tags) else str(t) for t in tags]


@dataclass
class ColumnSchema:
    name: str
    tags: List[str] = field(default_factory=list)
    dtype: str = "int64"              # "int64" | "float32"
    int_min: Optional[int] = None
    int_max: Optional[int] = None
    is_list: bool = True
    value_count_min: Optional[int] = None
    value_count_max: Optional[int] = None

    def __post_init__(self):
        self.tags = _norm_tags(self.tags)

    @classmethod
    def create_categorical(cls, name, num_items, tags=None, is_list=True, min_len=None, max_len=None):
        tags = _norm_tags(tags) + ["categorical"] + (["list"] if is_list else [])
        return cls(name, list(dict.fromkeys(tags)), "int64", 0 if num_items else None, num_items, is_list, min_len,
                   max_len)

    @classmethod
    def create_continuous(cls, name, tags=None, is_list=True, min_len=None, max_len=None):
        tags = _norm_tags(tags) + ["continuous"] + (["list"] if is_list else [])
        return cls(name, list(dict.fromkeys(tags)), "float32", None, None, is_list, min_len, max_len)


class Schema:
    def __init__(self, columns: Optional[Iterable[ColumnSchema]] = None):
        self.feature: List[ColumnSchema] = list(columns or [])

    # selection API used by from_schema (reference: schema.py:355-460)
    def select_by_tag(self, tags) -> "Schema":
        want = set(_norm_tags(tags))
        return Schema([c for c in self.feature if want & set(c.tags)])

    def select_by_name(self, names) -> "Schema":
        if isinstance(names, str):
            names = [names]
        names = set(names)
        return Schema([c for c in self.feature if c.name in names])

    def remove_by_tag(self, tags) -> "Schema":
        drop = set(_norm_tags(tags))
        return Schema([c for c in self.feature if not (drop & set(c.tags))])

    @property
    def column_names(self) -> List[str]:
        return [c.name for c in self.feature]

    @property
    def item_id_column_name(self) -> str:
        cols = self.select_by_tag(Tags.ITEM_ID).column_names
        if not cols:
            raise ValueError("There is no column tagged as item id.")
        return cols[0]

    def __iter__(self):
        return iter(self.feature)

    def __len__(self):
        return len(self.feature)

    def __bool__(self):
        return len(self.feature) > 0

    def __add__(self, other: "Schema") -> "Schema":
        return Schema(self.feature + other.feature)

    def categorical_cardinalities(self) -> Dict[str, int]:
        """schema.py:541-550: rows of the table = int_domain.max + 1."""
        out = {}
        for c in self.feature:
            if c.int_max is not None and c.int_max > 0:
                out[c.name] = int(c.int_max) + 1
        return out

    @classmethod
    def from_json(cls, path_or_str: str) -> "Schema":
        """Parse a TF-metadata style JSON schema (the format of the reference's
        transformers4rec/data/testing/schema.json)."""
        try:
            with open(path_or_str) as f:
                doc = json.load(f)
        except (OSError, ValueError):
            doc = json.loads(path_or_str)
        cols = []
        for f in doc.get("feature", []):
            tags = list(f.get("annotation", {}).get("tag", []))
            vc = f.get("valueCount") or f.get("value_count") or {}
            dom = f.get("intDomain") or f.get("int_domain")
            is_float = f.get("type", "INT") == "FLOAT"
            cols.append(ColumnSchema(
                name=f["name"], tags=tags, dtype="float32" if is_float else "int64",
                int_min=int(dom.get("min", 0)) if dom else None,
                int_max=int(dom["max"]) if dom and "max" in dom else None,
                is_list=bool(vc) or "list" in tags,
                value_count_min=int(vc["min"]) if "min" in vc else None,
                value_count_max=int(vc["max"]) if "max" in vc else None))
        return cls(cols)


def categorical_cardinalities(schema: Schema) -> Dict[str, int]:
    return schema.categorical_cardinalities()
