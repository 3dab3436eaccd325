// The two plain structs of the reference's src/common/vision_structs.h that the operator API uses: VRect
// (:122-133, the crop ROI) and VPoint (:6-34, centre of get_rotation_matrix_2D).  The remaining application
// types of that header (eye/gesture/state records) are not part of the preprocessing path and are not provided.
#ifndef VISION_STRUCTS_H
#define VISION_STRUCTS_H

namespace vision {

class VPoint {
public:
    VPoint() : x(0.0F), y(0.0F) {}
    VPoint(float _x, float _y) : x(_x), y(_y) {}
    float x;
    float y;
};

struct VRect {
    float left;
    float top;
    float right;
    float bottom;
    VRect(float _left, float _top, float _right, float _bottom) : left(_left), top(_top), right(_right), bottom(_bottom) {}
    void set(float left, float top, float right, float bottom);
    float width() const;    // right - left
    float height() const;   // bottom - top
    bool contains(float x, float y);
};

}  // namespace vision

#endif  // VISION_STRUCTS_H
