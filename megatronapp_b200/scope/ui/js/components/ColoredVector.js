// One compressed activation vector as a strip of coloured cells (reference components/ColoredVector.vue).
import {diverging} from "../draw.js";
export function coloredVector(values, max, name = "") {
  const row = document.createElement("div"); row.className = "vec";
  const lab = document.createElement("span"); lab.className = "name"; lab.textContent = name; lab.title = name; row.appendChild(lab);
  const cv = document.createElement("canvas"); cv.width = Math.max(1, values.length); cv.height = 1; row.appendChild(cv);
  const ctx = cv.getContext("2d"), img = ctx.createImageData(cv.width, 1), mx = max || Math.max(1e-12, ...values.map(Math.abs));
  values.forEach((v, i) => { const c = diverging(v / mx); img.data[4 * i] = c[0]; img.data[4 * i + 1] = c[1]; img.data[4 * i + 2] = c[2]; img.data[4 * i + 3] = 255; });
  ctx.putImageData(img, 0, 0); cv.title = values.slice(0, 8).map(v => (+v).toPrecision(3)).join(", ") + (values.length > 8 ? ", ..." : "");
  return row;
}
