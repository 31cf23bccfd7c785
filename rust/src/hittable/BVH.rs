//! BVHNode (reference: hittable/BVH.rs:7-90).  The device traverses its own trees (a binned-SAH binary tree and the
//! eight-wide compressed tree derived from it); this node only reproduces the SHAPE of the reference's median-split tree,
//! because its left-then-right visiting order (BVH.rs:71-79) is what ranks primitives for exact ties (gs_quad.order).
use crate::camera::scene_builder::{SceneBuilder, Xform};
use crate::hittable::hittable::{Hittable, HittableList};
use crate::AABB::AABB;

pub struct BVHNode {
    left: Box<dyn Hittable>,
    right: Option<Box<dyn Hittable>>,
    bbox: AABB,
}

impl BVHNode {
    pub fn from_list(list: HittableList) -> Box<Self> {
        BVHNode::construct_tree(list.objects)
    }

    /// one object: a node with a single child; two: in the given order; more: stable sort by the lower bound along the longest
    /// axis of the hull, split at len / 2 (BVH.rs:18-65)
    pub fn construct_tree(mut objects: Vec<Box<dyn Hittable>>) -> Box<BVHNode> {
        assert!(!objects.is_empty(), "BVHNode over an empty list (the reference recurses forever here)");
        if objects.len() == 1 {
            let only = objects.pop().unwrap();
            return Box::new(BVHNode { bbox: only.bounding_box(), left: only, right: None });
        }
        if objects.len() == 2 {
            let second = objects.pop().unwrap();
            let first = objects.pop().unwrap();
            let bbox = AABB::from_AABB_pair(first.bounding_box(), second.bounding_box());
            return Box::new(BVHNode { bbox, left: first, right: Some(second) });
        }
        let bbox = objects.iter().fold(AABB::EMPTY, |acc, o| AABB::from_AABB_pair(acc, o.bounding_box()));
        let axis = bbox.longest_axis();
        objects.sort_by(|a, b| {
            a.bounding_box().axis(axis).min.partial_cmp(&b.bounding_box().axis(axis).min).expect("NaN bounding box")
        });
        let upper = objects.split_off(objects.len() / 2);
        let left = BVHNode::construct_tree(objects);
        let right = BVHNode::construct_tree(upper);
        Box::new(BVHNode { left, right: Some(right), bbox })
    }
}

impl Hittable for BVHNode {
    fn bounding_box(&self) -> AABB {
        self.bbox
    }
    fn flatten(&self, xf: &Xform, out: &mut SceneBuilder) {
        self.left.flatten(xf, out);
        if let Some(r) = &self.right {
            r.flatten(xf, out);
        }
    }
}
