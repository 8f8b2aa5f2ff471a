// MegaScope front-end entry point: mounts the components (reference transformer-visualize/src/AppContent.vue layout:
// control column on the left, token strip + selectors on top, a 2-column grid of views).
import {ControlPanel} from "./components/ControlPanel.js";
import {TokenStrip} from "./components/TokenStrip.js";
import {Selectors} from "./components/Selectors.js";
import {VectorView} from "./components/VectorViews.js";
import {AttentionMatrix} from "./components/AttentionMatrix.js";
import {OutputProbs} from "./components/OutputProbs.js";
import {PCAPlot} from "./components/PCAPlot.js";
import {h} from "./components/base.js";

const main = document.querySelector("main");
new ControlPanel().mount(main);
const right = h("section"); main.appendChild(right);
new TokenStrip().mount(right); new Selectors().mount(right);
const grid = h("div", {class:"grid2"}); right.appendChild(grid);
new VectorView(1, "QKV projections (token x pixel)", true).mount(grid);
new AttentionMatrix().mount(grid);
new VectorView(4, "MLP fc1").mount(grid);
new VectorView(5, "MLP fc2").mount(grid);
new OutputProbs().mount(grid);
new PCAPlot().mount(grid);
