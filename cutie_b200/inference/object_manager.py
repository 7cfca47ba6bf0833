"""Object-id bookkeeping (host side, no kernels) -- cutie/inference/object_manager.py:7-149,
cutie/inference/object_info.py:1-24.

Object ids are what the user supplies and never change; *tmp ids* are 1-based positions of the
live objects along the object axis of every [B, K, ...] tensor and are re-packed on deletion.
"""
from typing import Dict, List, Union

import torch


class ObjectInfo:
    def __init__(self, id: int):
        self.id = id
        self.poke_count = 0        # consecutive missed detections (used by BURST-style drivers)

    def poke(self) -> None:
        self.poke_count += 1

    def unpoke(self) -> None:
        self.poke_count = 0

    def __hash__(self):
        return hash(self.id)

    def __eq__(self, other):
        return self.id == (other if type(other) == int else other.id)

    def __repr__(self):
        return f'(ID: {self.id})'


class ObjectManager:
    def __init__(self):
        self.obj_to_tmp_id: Dict[ObjectInfo, int] = {}
        self.tmp_id_to_obj: Dict[int, ObjectInfo] = {}
        self.obj_id_to_obj: Dict[int, ObjectInfo] = {}
        self.all_historical_object_ids: List[int] = []

    def _reindex(self) -> None:
        self.obj_id_to_obj = {o.id: o for o in self.obj_to_tmp_id}

    def add_new_objects(self, objects: Union[List[ObjectInfo], ObjectInfo, List[int]]) -> (List[int], List[int]):
        if not isinstance(objects, list):
            objects = [objects]
        tmp_ids, obj_ids = [], []
        for o in objects:
            oid = o if isinstance(o, int) else o.id
            known = self.obj_id_to_obj.get(oid)
            if known is None:
                known = ObjectInfo(id=oid)
                slot = len(self.obj_to_tmp_id) + 1
                self.obj_to_tmp_id[known] = slot
                self.tmp_id_to_obj[slot] = known
                self.obj_id_to_obj[oid] = known
                self.all_historical_object_ids.append(oid)
            tmp_ids.append(self.obj_to_tmp_id[known])
            obj_ids.append(oid)
        self._reindex()
        assert tmp_ids == sorted(tmp_ids)
        return tmp_ids, obj_ids

    def delete_objects(self, obj_ids_to_remove: Union[int, List[int]]) -> None:
        if isinstance(obj_ids_to_remove, int):
            obj_ids_to_remove = [obj_ids_to_remove]
        survivors = [self.tmp_id_to_obj[t] for t in range(1, len(self.tmp_id_to_obj) + 1)
                     if self.tmp_id_to_obj[t].id not in obj_ids_to_remove]
        self.obj_to_tmp_id = {o: i for i, o in enumerate(survivors, start=1)}
        self.tmp_id_to_obj = {i: o for i, o in enumerate(survivors, start=1)}
        self._reindex()

    def purge_inactive_objects(self, max_missed_detection_count: int) -> (bool, List[int], List[int]):
        gone = [o for o in self.obj_to_tmp_id if o.poke_count > max_missed_detection_count]
        kept = [o for o in self.obj_to_tmp_id if o.poke_count <= max_missed_detection_count]
        tmp_keep = [self.obj_to_tmp_id[o] for o in kept]
        if gone:
            self.delete_objects([o.id for o in gone])
        return len(gone) > 0, tmp_keep, [o.id for o in kept]

    def tmp_to_obj_cls(self, mask) -> torch.Tensor:
        """tmp-id class map -> object-id class map (object_manager.py:99-104) as one table lookup.  The table lives
        on the mask's device and is rebuilt only when the object set changes: boolean-mask assignment per object
        (or a per-frame pageable H2D copy) would stall the host on the GPU every frame."""
        return self.tmp_to_obj_lut(mask.device, mask.dtype)[mask]

    def tmp_to_obj_lut(self, device, dtype=torch.int64) -> torch.Tensor:
        """[1 + num_obj] table tmp id -> object id (0 -> 0), cached on `device` until the object set changes."""
        sig = (tuple((t, o.id) for t, o in self.tmp_id_to_obj.items()), device, dtype)
        if getattr(self, '_lut_sig', None) != sig:
            lut = torch.zeros(len(self.tmp_id_to_obj) + 1, dtype=dtype)
            for tmp_id, obj in self.tmp_id_to_obj.items():
                lut[tmp_id] = obj.id
            self._lut, self._lut_sig = lut.to(device), sig
        return self._lut

    def get_tmp_to_obj_mapping(self) -> Dict[int, ObjectInfo]:
        return {obj.id: tmp_id for obj, tmp_id in self.tmp_id_to_obj.items()}

    def realize_dict(self, obj_dict, dim=1) -> torch.Tensor:
        """{obj id: tensor} -> one tensor stacked along `dim` in tmp-id order."""
        parts = []
        for _, obj in self.tmp_id_to_obj.items():
            if obj.id not in obj_dict:
                raise NotImplementedError
            parts.append(obj_dict[obj.id])
        return torch.stack(parts, dim=dim)

    def make_one_hot(self, cls_mask) -> torch.Tensor:
        planes = [cls_mask == obj.id for _, obj in self.tmp_id_to_obj.items()]
        if not planes:
            return torch.zeros((0, *cls_mask.shape), dtype=torch.bool, device=cls_mask.device)
        return torch.stack(planes, dim=0)

    @property
    def all_obj_ids(self) -> List[int]:
        return [o.id for o in self.obj_to_tmp_id]

    @property
    def num_obj(self) -> int:
        return len(self.obj_to_tmp_id)

    def has_all(self, objects: List[int]) -> bool:
        return all(o in self.obj_to_tmp_id for o in objects)

    def find_object_by_id(self, obj_id) -> ObjectInfo:
        return self.obj_id_to_obj[obj_id]

    def find_tmp_by_id(self, obj_id) -> int:
        return self.obj_to_tmp_id[self.obj_id_to_obj[obj_id]]
